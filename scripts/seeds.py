"""Times extract_block on a few synthetic 128^3 blocks (argv: seeds...)."""
import sys, time, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
dev = torch.device('cuda:0')
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
seeds = [int(a) for a in sys.argv[1:]] or [500, 501, 502, 505]
out = []
for seed in seeds:
    f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
    ts = []
    for rep in range(4):
        torch.manual_seed(rep)
        torch.cuda.synchronize(); t = time.time()
        g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
        torch.cuda.synchronize(); ts.append((time.time() - t) * 1e3)
    out.append('%d: %.2f (kept %d)' % (seed, min(ts[1:]), m.numel()))
print(' | '.join(out), flush=True)
