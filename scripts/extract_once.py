"""Runs extract_block a few times on one synthetic 128^3 block (profiling driver)."""
import sys, time, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 500
f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
for rep in range(reps):
    torch.cuda.synchronize(); t = time.time()
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    torch.cuda.synchronize()
    print('rep', rep, 'ms %.2f' % ((time.time() - t) * 1e3), 'kept', m.numel(), flush=True)
