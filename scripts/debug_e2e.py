"""Why is the multi-stream end-to-end path slow?  H2D of the NeRF parameters from worker threads, alone and with the extract."""
import os, sys, time, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg
dev = torch.device("cuda:0")
occ, poses = pkg.synthetic.extract_scene(128, 50)
meta_host = pkg.synthetic.extract_meta(poses)
sgrid = pkg.SampleGrid(list(pkg.synthetic.AABB), 128)
fields = [pkg.synthetic.make_ngp_field(seed=500 + s) for s in (0, 1)]
host = [(f.mlp_base.params.detach().clone().pin_memory(), f.color_mlp.params.detach().clone().pin_memory()) for f in fields]
occ_p, poses_p = occ.to(torch.uint8).pin_memory(), poses.pin_memory()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bufs = [dict(fields=[pkg.synthetic.make_ngp_field(seed=950 + s).to(dev) for s in (0, 1)],
             occ=torch.empty_like(occ_p, device=dev), poses=torch.empty_like(poses_p, device=dev)) for _ in range(S)]
from importlib import import_module
slot = import_module("dreg-nerf_b200.nerf_regtr").engine_slot
pipe = pkg.PairPipeline(dev, streams=S)

def copy_only(i):
    sb = bufs[slot()]
    t0 = time.perf_counter()
    with torch.no_grad():
        for f, (hp, hc) in zip(sb["fields"], host):
            f.mlp_base.params.data.copy_(hp, non_blocking=True)
            f.color_mlp.params.data.copy_(hc, non_blocking=True)
        sb["occ"].copy_(occ_p, non_blocking=True)
        sb["poses"].copy_(poses_p, non_blocking=True)
    t1 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    return t1 - t0, time.perf_counter() - t1

def copy_and_extract(i):
    sb = bufs[slot()]
    with torch.no_grad():
        for f, (hp, hc) in zip(sb["fields"], host):
            f.mlp_base.params.data.copy_(hp, non_blocking=True)
            f.color_mlp.params.data.copy_(hc, non_blocking=True)
        sb["occ"].copy_(occ_p, non_blocking=True)
        sb["poses"].copy_(poses_p, non_blocking=True)
        meta = dict(meta_host, camera_poses=sb["poses"])
        g = [pkg.extract_block(f, sgrid, sb["occ"].bool(), meta, dev) for f in sb["fields"]]
    return g[0][1].numel()

def extract_only(i):
    sb = bufs[slot()]
    with torch.no_grad():
        meta = dict(meta_host, camera_poses=sb["poses"])
        g = [pkg.extract_block(f, sgrid, sb["occ"].bool(), meta, dev) for f in sb["fields"]]
    return g[0][1].numel()

for name, fn in (("copy only", copy_only), ("extract only", extract_only), ("copy + extract", copy_and_extract)):
    pipe.map(fn, range(2 * S))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = pipe.map(fn, range(16))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 16
    extra = ""
    if name == "copy only":
        extra = " enqueue %.2f ms, wait %.2f ms per pair; %.1f GB/s aggregate" % (
            1e3 * sum(o[0] for o in out) / 16, 1e3 * sum(o[1] for o in out) / 16, 103e6 / dt / 1e9)
    print("%d streams, %s: %.2f ms per pair%s" % (S, name, dt * 1e3, extra), flush=True)
