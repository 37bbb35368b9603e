cd /root/repo
python - <<'PY'
import sys, time, ctypes, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
lib = pkg.load_library()
dev = torch.device('cuda:0')
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
for seed in (500, 501, 502, 505):
    f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    torch.cuda.synchronize()
    st = (ctypes.c_ulonglong * 4)()
    lib.drb_march_stats(st, 1)
    torch.cuda.synchronize(); t = time.time()
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    torch.cuda.synchronize(); dt = (time.time() - t) * 1e3
    lib.drb_march_stats(st, 1)
    rays, skips, samples, rounds = [int(v) for v in st]
    pts = torch.rand(200000, 3, device=dev) * 3 - 1.5
    d = f.query_density(pts)
    print('seed', seed, 'kept', m.numel(), 'ms %.1f' % dt, 'rays %.2fM skips/ray %.1f samples/ray %.2f samples %.1fM warp-rounds %.2fM' % (rays/1e6, skips/max(rays,1), samples/max(rays,1), samples/1e6, rounds/1e6),
          'density>0.7 frac %.3f median %.3f' % (float((d > 0.7).float().mean()), float(d.median())), flush=True)
PY
