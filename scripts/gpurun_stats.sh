cd /root/repo
python - <<'PY'
import sys, time, ctypes, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
import os; lib = ctypes.CDLL(os.environ.get('DRB_LIB_PATH') or 'dreg-nerf_b200/libdregb200.so')
dev = torch.device('cuda:0')
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
for seed in (500, 501, 502, 503):
    f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    torch.cuda.synchronize()
    st = (ctypes.c_ulonglong * 4)()
    lib.drb_debug_march_stats(st, 1)
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    ts = []
    for rep in range(6):
        torch.cuda.synchronize(); t = time.time()
        g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
        torch.cuda.synchronize(); ts.append((time.time() - t) * 1e3)
    ts.sort(); dt = ts[len(ts)//2]
    print('   times', [round(v,1) for v in ts])
    lib.drb_debug_march_stats(st, 1)
    rays, skips, samples, iters = [int(v) / 2 for v in st]
    print('seed', seed, 'kept', m.numel(), 'ms', round(dt, 1), 'rays %.2fM skips/ray %.1f samples/ray %.2f warp-iters %.2fM samples/warp-iter %.1f' % (rays/1e6, skips/max(rays,1), samples/max(rays,1), iters/1e6, samples/max(iters,1)))
PY
