#!/usr/bin/env python
"""Benchmark of the DReg-NeRF registration hot path on B200 (see DESIGN.md "Measurement").

Metric (BASELINE.json): NeRF pairs/sec at a 128^3 grid.  Workload = configs[1]: single pair per
step, 128^3, fp32-grade arithmetic, synthetic random-weight network and synthetic pairs.

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W     # reference CPU arm (oracle port)

One JSON line on stdout (rank 0).  `value` = whole-job pairs/s with the inputs already resident in
HBM; `e2e` = the same metric through the public API from pinned HOST buffers (H2D of both grids and
masks and the D2H of the pose inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RES = 128
HOST_WAIT = "default (spin)"     # how host threads wait for the GPU (set by _blocking_sync)
DEFAULT_FULL_STREAMS = 4        # pairs in flight of the full stage (1 = one pair at a time, as round 1 ran it)
N_RESIDENT_PAIRS = 3        # distinct synthetic pairs cycled through the timed steps
FPN_FLOPS_PER_GRID_128 = 1384.0e9   # BASELINE.md section 2 (measured, 2*MAC)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.samples.append([f.strip() for f in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(s[0]) for s in self.samples if len(s) == 6 and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) == 6 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) == 6 and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _state(pkg, model):
    return pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0)


# ------------------------------------------------------------------------------------------------
def cpu_extract_seconds(pkg, cams, ray_sample=20000, like_fraction=0.125):
    """CPU time of ONE block's extract (oracle port: the reference's own extract is CUDA-only, tcnn +
    nerfacc, so no reference CPU implementation exists).  Density / colour over all candidate cells is
    timed in full (torch CPU ops).  The surface-field ray march runs in the C / OpenMP restatement
    (oracle/extract_c.c, all host threads) twice, each on a bounded sample scaled by its sample fraction:
      like_for_like - the rays OUR arm marches: only cells whose density passes the threshold, cameras in order
                      until one sees the point (identical masks; `like_fraction` of the candidate cells);
      as_reference  - every (camera, point) ray to the end, as sample_grid.py:245-318 marches them
                      (`ray_sample` rays).
    Returns (seconds per block with the like-for-like march, info dict with both)."""
    import torch
    from oracle import extract, extract_c, ngp
    from oracle.make_goldens import make_field
    occ, poses = pkg.synthetic.extract_scene(RES, cams)
    meta = pkg.synthetic.extract_meta(poses)
    _, ref = make_field(pkg, 500, 8.0)
    idx = torch.nonzero(occ.flatten())[:, 0]
    gen = torch.Generator().manual_seed(0)
    jitter = torch.rand(idx.numel(), 3, generator=gen)
    roi = list(pkg.synthetic.AABB)
    t0 = time.perf_counter()
    pts = extract.sample_points(idx, jitter, RES, roi)
    dens, feat = ngp.query_density(pts, ref["aabb"], ref["table"], ref["w1"], ref["w2"])
    ngp.query_rgb_mean(ngp.fixed_viewing_directions(), feat, ref["c1"], ref["c2"], ref["c3"])
    t_field = time.perf_counter() - t0
    cam_o = poses[:, :3, 3].contiguous()
    extract_c.load()
    # (1) like for like
    n_like = max(1, int(idx.numel() * like_fraction))
    pick = torch.randperm(idx.numel(), generator=gen)[:n_like]
    dense = (dens[pick] > 0.7)
    t0 = time.perf_counter()
    _, _, n_s_like = extract_c.surface_mask(pts[pick], cam_o, occ, RES, roi, roi, meta["render_step_size"], 0.5, ref,
                                            active=dense, all_rays=False)
    t_like_meas = time.perf_counter() - t0
    t_like = t_like_meas * (idx.numel() / n_like)
    # (2) as the reference marches
    total_rays = idx.numel() * cams
    n_pts = max(1, min(idx.numel(), ray_sample // max(cams, 1)))
    pick2 = torch.randperm(idx.numel(), generator=gen)[:n_pts]
    t0 = time.perf_counter()
    _, _, n_samples = extract_c.surface_mask(pts[pick2], cam_o, occ, RES, roi, roi, meta["render_step_size"], 0.5, ref,
                                             all_rays=True)
    t_all_meas = time.perf_counter() - t0
    t_rays = t_all_meas * (total_rays / (n_pts * cams))
    return t_field + t_like, {
        "field_s": t_field,
        "like_for_like": {"measured_s": t_like_meas, "cells_sampled": int(n_like), "cells_total": int(idx.numel()),
                          "dense_in_sample": int(dense.sum()), "density_samples_in_sample": int(n_s_like),
                          "block_s_extrapolated": t_like, "what": "dense cells only, cameras until one sees the point (the rays our arm marches)"},
        "as_reference": {"measured_s": t_all_meas, "rays_sampled": int(n_pts * cams), "rays_total": int(total_rays),
                         "density_samples_in_sample": int(n_samples), "block_s_extrapolated": t_rays,
                         "what": "every (camera, point) ray marched to its end (sample_grid.py:245-318)"},
        "measured_s_total": t_field + t_like_meas + t_all_meas,
        "marcher": "oracle/extract_c.c (C, OpenMP, %d threads)" % (os.cpu_count() or 1)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on all host threads.  Register
    half = oracle/regtr.py, bit-identical to the reference's modules (tests/golden).  Extract half = the
    oracle port (the reference has no CPU extract), bounded sample, see cpu_extract_seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import dreg_nerf_b200 as pkg   # synthetic inputs only; the CUDA path is not touched by this arm
    from oracle import regtr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = pkg.NeRFRegTr()
    sd = _state(pkg, model)
    del model
    data = pkg.synthetic.make_pair(res=RES, pair_id=0)
    if args.stage == "train":
        model = pkg.NeRFRegTr()
        model.load_state_dict(sd)
        model.correspondence_decoder.q_norm.requires_grad_(False)
        cb = cpu_train_baseline(pkg, model, steps=max(1, min(args.steps, 3)))
        line = {"impl": "reference", "metric": "nerf_pairs_per_sec_128cube", "value": cb["value"], "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "128^3 training step (fwd + loss + bwd + clip + AdamW), one pair per step (bounded sample of the batch)",
                           "stage": "train", "resolution": RES},
                "cpu_baseline": cb, "gpu_launches": 0,
                "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    full = args.stage == "full"
    sample = "1 pair, 128^3, full NeRFRegTr.forward (oracle port of the reference, fp32, torch CPU ops)"

    def full_step():
        with torch.no_grad():
            regtr.forward(sd, data, training=True)

    def fpn_once():
        with torch.no_grad():
            t0 = time.perf_counter()
            regtr.fpn3d(data["src_xyz_rgba"][:, 3:], sd, training=True)
            return time.perf_counter() - t0

    extract_s, extract_info = (0.0, None)
    if full:
        one_block, extract_info = cpu_extract_seconds(pkg, args.cams, ray_sample=200000, like_fraction=0.125)
        extract_s = 2.0 * one_block
    t0 = time.perf_counter()
    full_step()
    first = time.perf_counter() - t0
    budget = 200.0
    times = []
    if first * (args.steps + max(args.warmup - 1, 0)) <= budget:
        for _ in range(max(args.warmup - 1, 0)):
            full_step()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            full_step()
            times.append(time.perf_counter() - t0 + extract_s)
    else:
        # bounded sample: per step one grid's FPN; pair time = 2 x FPN + the non-FPN tail of one full pair
        tail = max(first - 2.0 * fpn_once(), 0.0)
        sample = ("per step: FPN of ONE 128^3 grid (half a pair's conv work); pair time = 2 x that + the "
                  "non-FPN tail (%.2f s) measured on one full pair" % tail)
        for _ in range(args.steps):
            times.append(2.0 * fpn_once() + tail + extract_s)
    if full:
        lf, ar = extract_info["like_for_like"], extract_info["as_reference"]
        sample += ("; + extract of 2 blocks by the oracle port: field queries in full (%.1f s per block), ray march MEASURED on "
                   "%d of %d candidate cells (%.1f s) and EXTRAPOLATED by the cell count to %.1f s per block, marching the "
                   "rays our arm marches; marching every ray as the reference does would take %.0f s per block "
                   "(extrapolated from %d of %d rays)"
                   % (extract_info["field_s"], lf["cells_sampled"], lf["cells_total"], lf["measured_s"],
                      lf["block_s_extrapolated"], ar["block_s_extrapolated"], ar["rays_sampled"], ar["rays_total"]))
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": "nerf_pairs_per_sec_128cube", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("single pair, 128^3 grid, full extract->register forward, fp32" if full else
                                "single pair, 128^3 grid, register forward (NeRFRegTr.forward), fp32"),
                   "stage": args.stage, "resolution": RES, "pairs_per_step": 1, "bn_mode": "batch statistics",
                   "extract": extract_info, "register_s": (total / len(times)) - extract_s},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    with open("/proc/self/maps") as fh:       # the reference arm must not touch the product library
        line["native_so_loaded"] = "libdregb200" in fh.read()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dreg_nerf_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback for our arm"
    global HOST_WAIT
    HOST_WAIT = _blocking_sync(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    model = pkg.NeRFRegTr(precision=args.precision)
    model.load_state_dict(_state(pkg, model))
    model = model.to(dev).train(True)       # batch-statistics BatchNorm, as eval_nerf_regtr.py runs it
    # rank r owns pairs r, r + world, ... of one list of distinct pairs (independent units, no data-path
    # collective); the per-pair SE(3) of every rank meet in one all-gather per step (sharding.gather_poses)
    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")
    my_ids = sharding.shard_pairs(N_RESIDENT_PAIRS * world, rank, world)
    full = args.stage == "full"

    def gather_pose(pose):
        if world > 1:
            return sharding.gather_poses(pose.reshape(1, 3, 4).contiguous(), world)
        return pose
    stage_ms = {"extract": 0.0, "register": 0.0, "n": 0}

    def pin_grid_dict(p):
        q = {}
        for k, v in p.items():
            if torch.is_tensor(v):
                if v.dim() == 5:     # keep the [X,Y,Z,7] storage order of voxel_grid.pt
                    store = v.permute(0, 3, 4, 2, 1).contiguous().pin_memory()
                    q[k] = store.permute(0, 4, 3, 1, 2)
                else:
                    q[k] = v.contiguous().pin_memory()
            else:
                q[k] = v
        return q

    if not full:
        host_pairs = [pkg.synthetic.make_pair(res=RES, pair_id=i) for i in my_ids]
        dev_pairs = [pkg.synthetic.to_device(p, dev) for p in host_pairs]
        pinned = [pin_grid_dict(p) for p in host_pairs]
        h2d_bytes = sum(v.numel() * v.element_size() for k, v in pinned[0].items()
                        if torch.is_tensor(v) and k != "pose")
        masked = [int(dev_pairs[0]["src_mask"].numel()), int(dev_pairs[0]["tgt_mask"].numel())]

        def step_resident(i):
            with torch.no_grad():
                out = model(dict(dev_pairs[i % N_RESIDENT_PAIRS]))
            gather_pose(out["pose"][-1])      # per-pair SE(3), 48 B per rank
            return out

        # e2e: double-buffered H2D on a side stream (the copy of step i+1 overlaps the compute of step i);
        # every timed step still pays one full H2D of its inputs and one D2H of its pose.
        side = torch.cuda.Stream()
        slots = [{k: (torch.empty_strided(v.shape, v.stride(), dtype=v.dtype, device=dev) if torch.is_tensor(v) else v)
                  for k, v in pinned[0].items()} for _ in range(2)]
        for sl in slots:        # masks differ in length between pairs: allocate the longest
            for k in ("src_mask", "tgt_mask"):
                sl[k] = torch.empty(max(p[k].numel() for p in pinned), dtype=torch.int64, device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        state = {"primed": False}

        def prefetch(i):
            sl, src = slots[i % 2], pinned[i % N_RESIDENT_PAIRS]
            side.wait_event(consumed[i % 2])
            with torch.cuda.stream(side):
                for k, v in src.items():
                    if torch.is_tensor(v):
                        (sl[k][:v.numel()] if k.endswith("_mask") else sl[k]).copy_(v, non_blocking=True)
                ready[i % 2].record(side)

        def step_e2e(_):
            i = state["n"] = state.get("n", -1) + 1          # running step index (warm-up and timed steps)
            if not state["primed"]:
                for e_ in consumed:
                    e_.record()
                prefetch(i)
                state["primed"] = True
            src = pinned[i % N_RESIDENT_PAIRS]
            torch.cuda.current_stream().wait_event(ready[i % 2])
            prefetch(i + 1)
            sl = slots[i % 2]
            data = dict(sl)
            data["src_mask"] = sl["src_mask"][:src["src_mask"].numel()]
            data["tgt_mask"] = sl["tgt_mask"][:src["tgt_mask"].numel()]
            with torch.no_grad():
                out = model(data)
                pose = gather_pose(out["pose"][-1])
                consumed[i % 2].record()
                return pose.cpu()
    else:
        # extract inputs: two random-weight NeRF blocks per pair, a shell occupancy, a ring of cameras
        occ, poses = pkg.synthetic.extract_scene(RES, args.cams)
        meta_host = pkg.synthetic.extract_meta(poses)
        meta = dict(meta_host, camera_poses=poses.to(dev))
        occ_dev = occ.to(dev)
        sgrid = pkg.SampleGrid(list(pkg.synthetic.AABB), RES)
        host_fields, dev_fields = [], []
        # Weak scaling needs the SAME work per GPU whatever N: the cost of a block's ray march depends on its random
        # field (7 M ... 68 M density samples between seeds), so every rank extracts the same N_RESIDENT_PAIRS pairs of
        # fields - but with its own sampling jitter (a per-rank device generator seed below), so no two ranks see the
        # same points, grids, masks or poses: distinct pairs of equal cost.
        for i in range(N_RESIDENT_PAIRS):
            pair = [pkg.synthetic.make_ngp_field(seed=500 + 2 * i + side) for side in (0, 1)]
            host_fields.append([(f.mlp_base.params.detach().clone().pin_memory(),
                                 f.color_mlp.params.detach().clone().pin_memory()) for f in pair])
            dev_fields.append([f.to(dev) for f in pair])
        occ_pinned, poses_pinned = occ.to(torch.uint8).pin_memory(), poses.pin_memory()
        h2d_bytes = (sum(a.numel() * 4 + b.numel() * 4 for a, b in host_fields[0]) + occ_pinned.numel()
                     + poses_pinned.numel() * 4)
        masked = [0, 0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        model.reserve_mask_capacity(int(occ.sum()))      # no engine is rebuilt mid-run when a block keeps more cells

        def extract_and_register(fields, occ_d, meta_d, timed_stages, gather=True):
            grids = []
            if timed_stages:
                ev[0].record()
            for f in fields:
                grids.append(pkg.extract_block(f, sgrid, occ_d, meta_d, dev))
            if timed_stages:
                ev[1].record()
            data = {"src_xyz_rgba": grids[0][0].permute(3, 2, 0, 1).unsqueeze(0), "src_mask": grids[0][1],
                    "tgt_xyz_rgba": grids[1][0].permute(3, 2, 0, 1).unsqueeze(0), "tgt_mask": grids[1][1]}
            masked[0], masked[1] = int(grids[0][1].numel()), int(grids[1][1].numel())
            out = model(data)
            if timed_stages:
                ev[2].record()
            return gather_pose(out["pose"][-1]) if gather else out["pose"][-1]

        # several pairs in flight (--streams > 1, pipeline.PairPipeline): the same work per pair on a worker's own
        # stream / engine; the per-pair all-gather stays on the caller's thread (one communicator, one thread)
        from importlib import import_module as _imp
        _slot = _imp("dreg-nerf_b200.nerf_regtr").engine_slot
        multi_bufs = [dict(fields=[pkg.synthetic.make_ngp_field(seed=950 + side_).to(dev) for side_ in (0, 1)],
                           occ=torch.empty_like(occ.to(torch.uint8), device=dev), poses=torch.empty_like(poses, device=dev))
                      for _ in range(args.streams if args.streams > 1 else 0)]

        def pair_resident(i):
            with torch.no_grad():
                return extract_and_register(dev_fields[i % N_RESIDENT_PAIRS], occ_dev, meta, False, gather=False)

        def pair_e2e(i):
            sb = multi_bufs[_slot()]
            with torch.no_grad():
                for f, (hp, hc) in zip(sb["fields"], host_fields[i % N_RESIDENT_PAIRS]):
                    f.mlp_base.params.data.copy_(hp, non_blocking=True)
                    f.color_mlp.params.data.copy_(hc, non_blocking=True)
                sb["occ"].copy_(occ_pinned, non_blocking=True)
                sb["poses"].copy_(poses_pinned, non_blocking=True)
                meta_d = dict(meta_host, camera_poses=sb["poses"])
                return extract_and_register(sb["fields"], sb["occ"].bool(), meta_d, False, gather=False)

        def step_resident(i, timed_stages=False):
            with torch.no_grad():
                pose = extract_and_register(dev_fields[i % N_RESIDENT_PAIRS], occ_dev, meta, timed_stages)
            if timed_stages:
                torch.cuda.synchronize()
                stage_ms["extract"] += ev[0].elapsed_time(ev[1])
                stage_ms["register"] += ev[1].elapsed_time(ev[2])
                stage_ms["n"] += 1
            return pose

        # e2e: the NeRF parameters of step i+1 are copied on a side stream while step i computes
        side = torch.cuda.Stream()
        slot_fields = [[pkg.synthetic.make_ngp_field(seed=900 + s_ * 2 + side_).to(dev) for side_ in (0, 1)]
                       for s_ in range(2)]
        slot_occ = [torch.empty_like(occ_pinned, device=dev) for _ in range(2)]
        slot_poses = [torch.empty_like(poses_pinned, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        state = {"primed": False}

        def prefetch(i):
            j, sl = i % N_RESIDENT_PAIRS, i % 2
            side.wait_event(consumed[sl])
            with torch.cuda.stream(side), torch.no_grad():
                for f, (hp, hc) in zip(slot_fields[sl], host_fields[j]):
                    f.mlp_base.params.data.copy_(hp, non_blocking=True)
                    f.color_mlp.params.data.copy_(hc, non_blocking=True)
                slot_occ[sl].copy_(occ_pinned, non_blocking=True)
                slot_poses[sl].copy_(poses_pinned, non_blocking=True)
                ready[sl].record(side)

        def step_e2e(_):
            i = state["n"] = state.get("n", -1) + 1          # running step index (warm-up and timed steps)
            if not state["primed"]:
                for e_ in consumed:
                    e_.record()
                prefetch(i)
                state["primed"] = True
            sl = i % 2
            torch.cuda.current_stream().wait_event(ready[sl])
            prefetch(i + 1)
            with torch.no_grad():
                meta_d = dict(meta_host, camera_poses=slot_poses[sl])
                pose = extract_and_register(slot_fields[sl], slot_occ[sl].bool(), meta_d, False)
                consumed[sl].record()
                return pose.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes
    lib = pkg.load_library()
    surf_prof = {"ms": 0.0, "launches": 0, "stats": [0, 0, 0, 0]}

    def timed(fn, steps, profile=False):
        barrier()
        model.set_profile(profile)
        if profile:
            model.read_profile()
            if full:                      # surface-field kernel: events + work counters over the timed steps
                st = (ctypes.c_ulonglong * 4)()
                lib.drb_march_stats(st, 1)
                lib.drb_extract_set_profile(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = model.read_profile() if profile else None
        model.set_profile(False)
        if profile and full:
            ms_, n_ = ctypes.c_float(0.0), ctypes.c_int(0)
            lib.drb_extract_read_profile(ctypes.byref(ms_), ctypes.byref(n_))
            lib.drb_extract_set_profile(0)
            st = (ctypes.c_ulonglong * 4)()
            lib.drb_march_stats(st, 1)
            surf_prof.update(ms=ms_.value, launches=n_.value, stats=[int(v) for v in st])
        return float(t.item()), model.launch_count() - l0, prof

    torch.manual_seed(1000 + rank)      # per-rank sampling jitter from here on (make_ngp_field reseeds while it builds)
    multi = full and args.streams > 1
    pipe = pkg.PairPipeline(dev, streams=args.streams) if multi else None

    def timed_multi(pair_fn, steps, to_host):
        # `steps` pairs, args.streams of them in flight; the per-pair SE(3) all-gather (and the D2H read of the end-to-end
        # path) follow on the caller's stream in pair order
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        e0.record()
        for pose in pipe.map(pair_fn, range(steps)):
            pose = gather_pose(pose)
            if to_host:
                pose.cpu()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), model.launch_count() - l0

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    if multi:
        timed_multi(pair_resident, max(args.warmup, 3) * args.streams, False)      # every slot's engine warm
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    profile_note = "CUDA events around the kernel, inside the timed region"
    if multi:
        ms, launches = timed_multi(pair_resident, args.steps, False)
        ms_1, _, prof = timed(step_resident, args.steps, profile=True)
        profile_note = ("CUDA events around the kernel in a second pass over the same steps on ONE stream (%.2f ms per step): in "
                        "the timed region %d pairs are in flight and the brackets would span other pairs' kernels"
                        % (ms_1 / args.steps, args.streams))
    else:
        ms, launches, prof = timed(step_resident, args.steps, profile=True)
        ms_1 = ms
    clocks = sampler.summary() if sampler else None
    if multi:
        timed_multi(pair_e2e, 2 * args.streams, True)
        ms_e2e, _ = timed_multi(pair_e2e, args.steps, True)
    else:
        for i in range(2):
            step_e2e(i)
        ms_e2e, _, _ = timed(step_e2e, args.steps)
    if full:
        for i in range(min(args.steps, 3)):          # untimed extra steps: per-stage split
            step_resident(i, timed_stages=True)

    if rank == 0:
        peaks = _peaks()
        value = world * args.steps / (ms / 1e3)
        e2e = world * args.steps / (ms_e2e / 1e3)
        ig_ms, ig_flops, ig_n = prof
        achieved = ig_flops / (ig_ms / 1e3) / 1e12 if ig_ms > 0 else 0.0
        mma_factor = 3 if args.precision == "fp32" else 1
        peak = peaks["bf16_tflops_sustained"]
        ns, nt = model.last_token_counts
        line = {
            "metric": "nerf_pairs_per_sec_128cube", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": ("single pair per GPU, 128^3 grid, full extract->register forward on %dxB200, %s" if full else
                                    "single pair per GPU, 128^3 grid, register forward (NeRFRegTr.forward) on %dxB200, %s")
                                   % (world, "fp32" if args.precision == "fp32" else "bf16"),
                       "stage": args.stage,
                       "arithmetic": ("fp32 storage / accumulation; GEMM products as fp16 hi+lo operand pairs (22-bit "
                                      "significands, 3 tensor-core MMAs per product)") if args.precision == "fp32"
                                     else "bf16 GEMM operands, fp32 accumulation",
                       "extract": ({"candidate_cells_per_block": int(occ.sum()), "cameras": args.cams,
                                    "render_step_size": meta_host["render_step_size"],
                                    "stage_ms": {k: (v / max(stage_ms["n"], 1)) for k, v in stage_ms.items() if k != "n"}}
                                   if full else None),
                       "resolution": RES, "pairs_per_step_per_gpu": 1, "masked_voxels": masked, "host_wait": HOST_WAIT,
                       "tokens": [ns, nt], "bn_mode": "batch statistics",
                       "l2": "working set per step (2 x 58.7 MB grids, 0.6 GB weight planes, >3 GB activations) exceeds the 126 MB L2",
                       "parallelism": ("%d pairs per step over %d GPU(s), %d resident per rank: " % (world, world, N_RESIDENT_PAIRS)
                                       + ("the same NeRF fields on every rank (the march cost of a block depends on its field: fixed "
                                          "per-GPU work), sampled with a per-rank jitter seed - distinct points, grids, masks and poses per "
                                          "rank" if full else "distinct synthetic pairs per rank (sharding.shard_pairs)")
                                       + ", no data-path collective, one NCCL all-gather of the per-pair SE(3) per step")},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": 48, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv3d/linear), all launches of the timed steps",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peaks["source"] + ", bf16 sustained",
                         "traffic": None, "launches": int(ig_n), "kernel_ms_per_step": ig_ms / args.steps,
                         "kernel_share_of_step": ig_ms / ms_1, "timing": profile_note,
                         "mma_flops_factor": mma_factor, "tensor_pipe_frac_est": mma_factor * achieved / peak},
        }
        if multi:
            line["config"]["streams"] = ("%d pairs in flight per GPU, one CUDA stream + engine each (pipeline.PairPipeline); "
                                         "one pair at a time: %.2f ms per pair" % (args.streams, ms_1 / args.steps))
        if full:
            # In the full path the surface-field ray marcher is the dominant kernel.  DRAM is idle (the 48 MB
            # table is L2 resident); the kernel is bound by the rate at which an SM can miss 32-byte sectors
            # of the hash table into L2 (random 8-byte gathers, 128 per density sample).  Reported against
            # the HBM roofline as the contract asks: `achieved` = algorithmic bytes (every input read once),
            # `issued_*` = the gather traffic the algorithm issues (samples x 128 corners x 8 B).
            n_cand = int(occ.sum())
            table_b = int(lib.drb_ngp_table_entries()) * 8
            n_l = max(surf_prof["launches"], 1)
            surf_ms = surf_prof["ms"] / n_l                                   # per launch, inside the timed steps
            alg_bytes = table_b + RES ** 3 + n_cand * (12 + 1 + 1) + args.cams * 12   # one launch
            hb = alg_bytes / (surf_ms / 1e3) / 1e9 if surf_ms > 0 else 0.0
            rays, skips, samples, rounds = surf_prof["stats"]
            issued = samples * 1024.0 / n_l
            line["roofline_register"] = line["roofline"]
            line["roofline"] = {
                "bound": "hbm", "kernel": "surface_mask_kernel (occupancy-grid ray marcher + fused hash-grid / tensor-core "
                                          "MLP density), the %d launches of the timed steps" % surf_prof["launches"],
                "achieved": hb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hb / peaks["hbm_gbs"],
                "peak_source": peaks["source"], "traffic": 50.1e6,
                "traffic_note": "dram read+write of one launch, ncu --set full (profiles/r02_ncu_surface_seed501_keys.txt)",
                "kernel_ms_per_launch": surf_ms, "launches": surf_prof["launches"],
                "kernel_ms_per_step": surf_prof["ms"] / args.steps,
                "kernel_share_of_step": surf_prof["ms"] / ms_1, "timing": profile_note,
                "per_launch": {"rays": rays / n_l, "skip_events": skips / n_l, "density_samples": samples / n_l},
                "issued_gather_bytes_per_launch": issued,
                "issued_gather_gbs": issued / (surf_ms / 1e3) / 1e9 if surf_ms > 0 else 0.0,
                "density_samples_per_s": samples / (surf_prof["ms"] / 1e3) if surf_prof["ms"] > 0 else 0.0,
                # The limit that binds: an SM's L1 retires one distinct 32-byte sector per load instruction per
                # clock whatever the occupancy or ILP (scripts/ubench/gather*.cu on this pool's B200: 287 G
                # sectors/s per chip, halved when the shared-memory carve-out leaves < ~64 KB of L1).  The kernel
                # presents ~59 sectors per density sample to the L1 tag stage (ncu lts__t_sectors / (1 - L1 hit rate) /
                # samples on the heaviest synthetic block, profiles/r02_ncu_surface_seed501_keys.txt; round 1: 62;
                # 15 levels x 8 corners = 120 before pair loads, cell-major dense levels and in-instruction
                # sharing between coherent lanes).
                "gather_roofline": {"l1_sectors_per_sample": 59,
                                    "achieved_g_sectors_per_s": 59.0 * samples / (surf_prof["ms"] / 1e3) / 1e9 if surf_prof["ms"] > 0 else 0.0,
                                    "peak_g_sectors_per_s": 287.0, "peak_source": "measured, scripts/ubench/gather.cu (profiles/r01_ubench_gather.txt)",
                                    "frac": 59.0 * samples / (surf_prof["ms"] / 1e3) / 287e9 if surf_prof["ms"] > 0 else 0.0},
                "note": "algorithmic bytes = hash table + occupancy grid + points + masks + cameras, each read once per "
                        "launch; the kernel is L1-miss / L2-gather bound (DRAM idle), not bandwidth bound"}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(pkg, model, args.stage, args.cams)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def _blocking_sync(dev_index):
    """Host threads that wait for the GPU (stream synchronisations of the down-sampler, of the mask read-back, of the
    worker threads of a stream pipeline) can sleep instead of spinning (DRB_BLOCKING_SYNC=1 ->
    cudaDeviceScheduleBlockingSync on the rank's device, set before torch creates the context): with 4 worker threads
    per rank a spinning wait keeps 4-5 cores per GPU busy, which matters on a box with few cores per GPU.  Measured
    on a 16-core box with one GPU: sleeping costs 3 % of the pipelined throughput (51.8 -> 50.1 pairs/s) and 11 % of
    the one-pair-at-a-time latency (wake-ups on the down-sampler's synchronisations), so the default stays the
    driver's (spin).  Returns what is in effect for the JSON line."""
    if os.environ.get("DRB_BLOCKING_SYNC", "0") != "1":
        return "default (spin)"
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so.12")
        if rt.cudaSetDevice(int(dev_index)) == 0 and rt.cudaSetDeviceFlags(4) == 0:      # cudaDeviceScheduleBlockingSync
            return "cudaDeviceScheduleBlockingSync"
    except OSError:
        pass
    return "default (spin; cudaSetDeviceFlags unavailable)"


def _setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback for our arm"
    global HOST_WAIT
    HOST_WAIT = _blocking_sync(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def _timed_with_profile(timed, step, pipe, pipe1, steps):
    """-> (ms of the timed region, launches, (kernel ms, flops, launches), note).  One stream: the tensor-core
    GEMM launches are event-bracketed inside the timed region itself.  Several streams: kernels of different pairs
    overlap, an event pair around one launch then also spans other pairs' kernels - the timed region runs without
    the brackets and the per-launch durations come from a second pass over the same steps on ONE stream."""
    if pipe.n == 1:
        ms, launches, prof = timed(step, steps, profile=True)
        return ms, launches, prof, {"ms": ms, "what": "CUDA events around every launch, inside the timed region"}
    ms, launches, _ = timed(step, steps)
    ms1, _, prof = timed(lambda i: step(i, pipe1), steps, profile=True)
    return ms, launches, prof, {"ms": ms1, "what": "CUDA events around every launch in a second pass over the same steps on ONE "
                                "stream (%.1f ms per step): in the timed region %d pairs are in flight and an event pair "
                                "around one launch would span other pairs' kernels" % (ms1 / steps, pipe.n)}


def _pin_pair(p):
    """Pinned host copy of one pair in the [X,Y,Z,7] storage order of voxel_grid.pt (the view the loader makes)."""
    import torch
    q = {}
    for k, v in p.items():
        if torch.is_tensor(v):
            if v.dim() == 5:
                store = v.permute(0, 3, 4, 2, 1).contiguous().pin_memory()
                q[k] = store.permute(0, 4, 3, 1, 2)
            else:
                q[k] = v.contiguous().pin_memory()
        else:
            q[k] = v
    return q


def _h2d_pair(pinned, dev):
    import torch
    return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in pinned.items()}


def _pair_bytes(p):
    import torch
    return sum(v.numel() * v.element_size() for k, v in p.items() if torch.is_tensor(v))


def run_train(args):
    """BASELINE.json configs[2]: one step = `--batch` pairs at 128^3 through forward + loss + backward (gradients
    accumulate over the pairs: the network is B = 1 by construction, nerf_regtr.py:144-147) + one fused
    clip_grad_norm_(0.1) + AdamW update (train_nerf_regtr.py:171-239).  bf16 GEMM operands by default."""
    import torch
    import torch.distributed as dist
    import dreg_nerf_b200 as pkg
    world, rank, local_rank, dev = _setup_dist()
    torch.manual_seed(0)
    precision = args.precision or "bf16"
    model = pkg.NeRFRegTr(precision=precision)
    model.load_state_dict(_state(pkg, model))
    model = model.to(dev).train(True)
    model.correspondence_decoder.q_norm.requires_grad_(False)      # unused by the forward (no gradient in the reference either)
    crit = pkg.RegistrationLoss().to(dev)
    params = [p for p in model.parameters() if p.requires_grad] + list(crit.parameters())
    opt = pkg.FusedAdamW(params, lr=1e-4, weight_decay=1e-4, max_grad_norm=0.1)
    B = args.batch
    n_res = min(B, 8)
    ids = [rank * n_res + i for i in range(n_res)]             # distinct pairs per rank
    host = [pkg.synthetic.make_pair(res=RES, pair_id=i) for i in ids]
    pinned = [_pin_pair(p) for p in host]
    resident = [pkg.synthetic.to_device(p, dev) for p in host]
    h2d = sum(_pair_bytes(p) for p in pinned) / n_res * B
    loss_buf = torch.zeros((), device=dev)

    def one_pair(data):
        out = model(data)
        loss, _ = crit(out, data["pose"])
        (loss / B).backward()
        return loss.detach()

    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")

    def grads_allreduce():
        # data-parallel training: gradients averaged over ranks, one flat all-reduce per step
        sharding.allreduce_gradients(params, world)

    # the pairs of a step are independent until the optimiser: `--streams` of them in flight (pipeline.PairPipeline)
    pipe = pkg.PairPipeline(dev, streams=args.streams)
    pipe1 = pkg.PairPipeline(dev, streams=1)
    model.reserve_mask_capacity(max(max(p["src_mask"].numel(), p["tgt_mask"].numel()) for p in host))

    def step_resident(i, pipe=pipe):
        opt.zero_grad(set_to_none=True)
        losses = pipe.map(lambda b: one_pair(dict(resident[(i * B + b) % n_res])), range(B))
        tot = torch.stack(losses).sum()
        grads_allreduce()
        opt.step()
        return tot

    def step_e2e(i):
        opt.zero_grad(set_to_none=True)
        losses = pipe.map(lambda b: one_pair(_h2d_pair(pinned[(i * B + b) % n_res], dev)), range(B))
        tot = torch.stack(losses).sum()
        grads_allreduce()
        opt.step()
        return float(tot.cpu())                       # D2H of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        model.set_profile(profile)
        if profile:
            model.read_profile()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = model.read_profile() if profile else None
        model.set_profile(False)
        return float(t.item()), model.launch_count() - l0, prof

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, prof, prof_note = _timed_with_profile(timed, step_resident, pipe, pipe1, args.steps)
    clocks = sampler.summary() if sampler else None
    step_e2e(0)
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    if rank == 0:
        peaks = _peaks()
        pairs = world * B * args.steps
        ig_ms, ig_flops, ig_n = prof
        achieved = ig_flops / (ig_ms / 1e3) / 1e12 if ig_ms > 0 else 0.0
        mma = 3 if precision == "fp32" else 1
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": "nerf_pairs_per_sec_128cube", "value": pairs / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "batch %d pairs per GPU, 128^3, %s training step (fwd + loss + bwd + clip + AdamW) on %dxB200"
                                   % (B, "bf16" if precision != "fp32" else "fp32-grade", world),
                       "stage": "train", "resolution": RES, "pairs_per_step_per_gpu": B,
                       "streams": "%d pairs in flight per GPU, one CUDA stream + engine each (pipeline.PairPipeline)" % pipe.n,
                       "distinct_pairs_resident_per_gpu": n_res, "bn_mode": "batch statistics",
                       "loss": "RegistrationLoss: InfoNCE (0.1) + robust-L1 correspondence both directions (1.0), synthetic GT pose",
                       "optimizer": "FusedAdamW lr 1e-4 wd 1e-4, clip_grad_norm 0.1 (train_nerf_regtr.py:96-102,232-235)",
                       "l2": "per-pair working set (2 x 58.7 MB grids, >4 GB of saved activations) exceeds the 126 MB L2",
                       "parallelism": "data parallel over %d GPU(s): distinct pairs per rank, one flat gradient all-reduce per step" % world},
            "e2e": {"value": pairs / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "igemm_kernel + wgrad_kernel (tcgen05 conv / linear forward, data gradient, weight gradient), all launches of the timed steps",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peaks["source"] + ", bf16 sustained", "traffic": None, "launches": int(ig_n),
                         "kernel_ms_per_step": ig_ms / args.steps, "kernel_share_of_step": ig_ms / prof_note["ms"],
                         "timing": prof_note["what"],
                         "mma_flops_factor": mma, "tensor_pipe_frac_est": mma * achieved / peak},
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_train_baseline(pkg, model)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_train_baseline(pkg, model, steps=1):
    """One pair through the oracle port's forward + autograd backward + torch AdamW on the host cores."""
    import torch
    from oracle import regtr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    names = [k for k, p in model.named_parameters() if p.requires_grad]
    leaf = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in sd.items()}
    crit = pkg.RegistrationLoss()
    opt = torch.optim.AdamW([leaf[k] for k in names] + list(crit.parameters()), lr=1e-4, weight_decay=1e-4)
    data = pkg.synthetic.make_pair(res=RES, pair_id=0)
    t0 = time.perf_counter()
    for _ in range(steps):
        opt.zero_grad()
        loss, _ = crit(regtr.forward(leaf, data, training=True), data["pose"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in leaf.values() if p.requires_grad], 0.1)
        opt.step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": 1.0 / dt, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d pair(s), 128^3: oracle/regtr.py forward + torch autograd backward + clip + torch AdamW, fp32, %d threads (%.1f s per pair)"
                      % (steps, cores, dt)}


def run_batch(args):
    """BASELINE.json configs[3]: a list of DISTINCT pairs sharded over the GPUs (sharding.shard_pairs: rank r takes
    pairs r, r + N, ...), bf16 inference, ONE all-gather of the per-pair SE(3) per batch (sharding.gather_poses).
    --pairs-per-gpu P: weak scaling (P x N pairs); --total-pairs T: strong scaling (T pairs whatever N)."""
    import torch
    import torch.distributed as dist
    import dreg_nerf_b200 as pkg
    from importlib import import_module
    sharding = import_module("dreg-nerf_b200.sharding")
    world, rank, local_rank, dev = _setup_dist()
    torch.manual_seed(0)
    precision = args.precision or "bf16"
    model = pkg.NeRFRegTr(precision=precision)
    model.load_state_dict(_state(pkg, model))
    model = model.to(dev).train(True)      # batch-statistics BatchNorm, as eval_nerf_regtr.py runs it
    res = args.res
    if args.max_tokens:
        model.set_max_tokens(args.max_tokens)     # configs[4]: ~8k tokens per cloud instead of the reference's 1500
    model.tc_attention = args.attention == "tc"
    strong = args.total_pairs > 0
    n_pairs = args.total_pairs if strong else args.pairs_per_gpu * world
    mine = sharding.shard_pairs(n_pairs, rank, world)
    n_res = min(len(mine), 32)             # resident distinct pairs per rank (cycled beyond that)
    host = [pkg.synthetic.make_pair(res=res, pair_id=pid) for pid in mine[:n_res]]
    pinned = [_pin_pair(p) for p in host]
    resident = [pkg.synthetic.to_device(p, dev) for p in host]
    h2d = sum(_pair_bytes(p) for p in pinned) / max(n_res, 1) * len(mine)

    pipe = pkg.PairPipeline(dev, streams=args.streams)     # `--streams` pairs in flight, one CUDA stream + engine each
    model.reserve_mask_capacity(max(max(p["src_mask"].numel(), p["tgt_mask"].numel()) for p in host))
    tokens = {}

    def one(fetch, j):
        pose = model(fetch(j))["pose"][-1, 0]
        tokens[j] = model.last_token_counts
        return pose

    pipe1 = pkg.PairPipeline(dev, streams=1)

    def batch(fetch, pipe=pipe):
        with torch.no_grad():
            local = torch.stack(pipe.map(lambda j: one(fetch, j), range(len(mine))))
        return sharding.gather_poses(local, n_pairs)           # [n_pairs, 3, 4] on every rank

    def step_resident(_, pipe=pipe):
        return batch(lambda j: dict(resident[j % n_res]), pipe)

    def step_e2e(_):
        return batch(lambda j: _h2d_pair(pinned[j % n_res], dev)).cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        model.set_profile(profile)
        if profile:
            model.read_profile()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = model.read_profile() if profile else None
        model.set_profile(False)
        return float(t.item()), model.launch_count() - l0, prof

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches, prof, prof_note = _timed_with_profile(timed, step_resident, pipe, pipe1, args.steps)
    clocks = sampler.summary() if sampler else None
    step_e2e(0)
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    if rank == 0:
        peaks = _peaks()
        total = n_pairs * args.steps
        ig_ms, ig_flops, ig_n = prof
        achieved = ig_flops / (ig_ms / 1e3) / 1e12 if ig_ms > 0 else 0.0
        mma = 3 if precision == "fp32" else 1
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": "nerf_pairs_per_sec_%dcube" % res, "value": total / (ms / 1e3), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "batch %d pairs sharded across %dxB200, %d^3, %s, register forward, one all-gather of SE(3) per batch"
                                   % (n_pairs, world, res, "bf16" if precision != "fp32" else "fp32-grade"),
                       "stage": "batch", "resolution": res, "tokens": list(tokens.get(0, model.last_token_counts)),
                       "streams": "%d pairs in flight per GPU, one CUDA stream + engine each (pipeline.PairPipeline)" % pipe.n,
                       "max_tokens": int(model.max_tokens), "attention": "tcgen05 (attention.cu)" if model.tc_attention else "mma.sync (transformer.cu)", "pairs_per_step": n_pairs, "pairs_per_step_per_gpu": len(mine),
                       "distinct_pairs_resident_per_gpu": n_res, "bn_mode": "batch statistics",
                       "l2": "per-pair working set (2 x 58.7 MB grids, >3 GB activations) exceeds the 126 MB L2",
                       "parallelism": "sharding.shard_pairs (round robin), no data-path collective, one NCCL all-gather of [%d,3,4] per batch"
                                      % ((n_pairs + world - 1) // world)},
            "e2e": {"value": total / (ms_e2e / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": n_pairs * 48, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv3d/linear), all launches of the timed steps",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peaks["source"] + ", bf16 sustained", "traffic": None, "launches": int(ig_n),
                         "kernel_ms_per_step": ig_ms / args.steps, "kernel_share_of_step": ig_ms / prof_note["ms"],
                         "timing": prof_note["what"],
                         "mma_flops_factor": mma, "tensor_pipe_frac_est": mma * achieved / peak},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(pkg, model, stage, cams):
    """The oracle (port of the reference, bit-identical to it for the register half) timed on the host
    cores: one pair (+ the bounded extract sample for the full path)."""
    import torch
    from oracle import regtr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    data = pkg.synthetic.make_pair(res=RES, pair_id=0)
    t0 = time.perf_counter()
    with torch.no_grad():
        regtr.forward(sd, data, training=True)
    dt = time.perf_counter() - t0
    sample = "1 pair, 128^3, NeRFRegTr.forward, fp32 torch CPU ops, %d threads, single cold run (%.1f s)" % (cores, dt)
    if stage == "full":
        one_block, info = cpu_extract_seconds(pkg, cams, ray_sample=100000, like_fraction=0.125)
        dt += 2.0 * one_block
        lf, ar = info["like_for_like"], info["as_reference"]
        sample += ("; + extract of 2 blocks by the oracle port (%.1f s): field queries in full, ray march measured on %d of %d "
                   "candidate cells (%.1f s) and extrapolated by the cell count, marching the rays our arm marches (every "
                   "ray as the reference marches them: %.0f s per block, extrapolated from %d of %d rays)"
                   % (2.0 * one_block, lf["cells_sampled"], lf["cells_total"], lf["measured_s"],
                      ar["block_s_extrapolated"], ar["rays_sampled"], ar["rays_total"]))
    return {"value": 1.0 / dt, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"],
                    help="default: fp32-grade for full / register (configs[1]), bf16 for train / batch (configs[2], [3])")
    ap.add_argument("--batch", type=int, default=32, help="train: pairs per step per GPU")
    ap.add_argument("--pairs-per-gpu", type=int, default=32, help="batch: weak scaling, pairs per GPU per step")
    ap.add_argument("--total-pairs", type=int, default=0, help="batch: strong scaling, total pairs per step")
    ap.add_argument("--res", type=int, default=RES, help="batch: grid resolution (256 for configs[4])")
    ap.add_argument("--max-tokens", type=int, default=0, help="batch: cap of the down-sampler (tokens of the pair); 0 = the reference's 3000")
    ap.add_argument("--attention", default="tc", choices=["tc", "mma"], help="tc = tcgen05 FlashAttention-style kernel")
    ap.add_argument("--streams", type=int, default=0,
                    help="pairs in flight per GPU (pipeline.PairPipeline; 1 = the sequential loop); default 4 for train / "
                         "batch, 1 for full / register")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage", default="full", choices=["full", "register", "train", "batch"],
                    help="full = extract (2 NeRF blocks -> voxel grids) + register (configs[1], the default); register = "
                         "NeRFRegTr.forward only; train = configs[2] (fwd + bwd + AdamW); batch = configs[3] (sharded list of pairs)")
    ap.add_argument("--cams", type=int, default=50)
    args = ap.parse_args()
    if args.streams <= 0:
        args.streams = 4 if args.stage in ("train", "batch") else (DEFAULT_FULL_STREAMS if args.stage == "full" else 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.stage == "train":
        run_train(args)
    elif args.stage == "batch":
        run_batch(args)
    else:
        args.precision = args.precision or "fp32"
        run_ours(args)


if __name__ == "__main__":
    main()
