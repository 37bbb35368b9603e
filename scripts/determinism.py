"""Runs extract_block several times with the same jitter and compares the voxel masks bit for bit."""
import sys, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
dev = torch.device('cuda:0')
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
k = int(occ.sum())
out = {}
for seed in (501, 505):
    f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
    jitter = torch.rand((k, 3), generator=torch.Generator().manual_seed(3)).to(dev)
    masks = []
    for rep in range(5):
        g, m = pkg.extract_block(f, sg, occ_d, meta, dev, jitter=jitter)
        masks.append(m.cpu())
    same = all(torch.equal(masks[0], x) for x in masks)
    print('seed', seed, 'kept', [x.numel() for x in masks], 'all equal', same, flush=True)
    out[seed] = masks[0]
torch.save(out, sys.argv[1])
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    for seed in out:
        a, b = set(out[seed].tolist()), set(ref[seed].tolist())
        print('seed', seed, 'vs reference file: only here', sorted(a - b)[:5], 'only there', sorted(b - a)[:5])
