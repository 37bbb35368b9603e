"""256^3 plumbing check (BASELINE.json configs[4] shape): extract of one block + register forward."""
import sys, time, math, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
dev = torch.device('cuda:0')
res = 256
t = time.time()
occ, poses = pkg.synthetic.extract_scene(res, 20)
meta = dict(pkg.synthetic.extract_meta(poses), camera_poses=poses.to(dev))
sg = pkg.SampleGrid(list(pkg.synthetic.AABB), res)
f = pkg.synthetic.make_ngp_field(seed=500).to(dev)
occ_d = occ.to(dev)
print('candidate cells', int(occ.sum()), flush=True)
for rep in range(2):
    torch.cuda.synchronize(); t = time.time()
    g, m = pkg.extract_block(f, sg, occ_d, meta, dev)
    torch.cuda.synchronize()
    print('extract 256^3: %.1f ms kept %d' % ((time.time() - t) * 1e3, m.numel()), flush=True)
del g, m, occ_d
torch.manual_seed(0)
model = pkg.NeRFRegTr().to(dev).eval()
data = pkg.synthetic.to_device(pkg.synthetic.make_pair(res=res, pair_id=0), dev)
print('masked', data['src_mask'].numel(), data['tgt_mask'].numel(), flush=True)
with torch.no_grad():
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        out = model(data)
        torch.cuda.synchronize()
        print('register 256^3: %.1f ms tokens %s' % ((time.time() - t) * 1e3, model.last_token_counts), flush=True)
print('pose finite', bool(torch.isfinite(out['pose']).all()), 'mem GB %.1f' % (torch.cuda.max_memory_allocated() / 2**30))
