cd /root/repo
mkdir -p gpurun_out
nproc; uptime
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_extract_gpu.py -x -s 2>&1 | tail -7
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
dev = torch.device('cuda:0')
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
occ_d = occ.to(dev)
for seed in (500, 501, 502, 503, 504, 505):
    f = pkg.synthetic.make_ngp_field(seed=seed).to(dev)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t = time.time()
        g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
        torch.cuda.synchronize(); ts.append((time.time() - t) * 1e3)
    print('seed', seed, 'kept', m.numel(), 'ms', [round(v, 2) for v in ts], flush=True)
PY
