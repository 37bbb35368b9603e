cd /root/repo
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q -m gpu -p no:cacheprovider --timeout 300"
$T tests/test_kernels_gpu.py 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -30
import sys, torch
sys.path.insert(0, '.')
import dreg_nerf_b200 as pkg
from oracle import regtr
dev = torch.device('cuda:0')
def rel(a,b):
    a,b=a.detach().double().cpu(), b.double(); return ((a-b).abs().max()/(b.abs().max()+1e-30)).item()
for res, train in [(32, False), (64, True), (128, True)]:
    for gain in [4.0, 8.0]:
        torch.manual_seed(0)
        model = pkg.NeRFRegTr(); sd = pkg.synthetic.seeded_state_dict(model, 0, gain); model.load_state_dict(sd)
        model = model.to(dev).train(train)
        data = pkg.synthetic.make_pair(res=res, pair_id=0)
        with torch.no_grad():
            out = model(pkg.synthetic.to_device(data, dev)); ref = regtr.forward(sd, data, training=train)
        e = {k: rel(out[k][0], ref[k][0]) for k in ('src_feats','src_kp_warped','tgt_kp_warped','src_overlap')}
        e['pose'] = rel(out['pose'], ref['pose'])
        print(res, train, gain, {k: '%.1e'%v for k,v in e.items()}, 'tokens', model.last_token_counts, flush=True)
PY
echo ==== bench
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
