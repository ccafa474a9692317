#!/usr/bin/env python
"""Headline benchmark: SE3ET-I 3DMatch-shaped inference, synthetic pairs, pairs/sec (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path on the host cores

A step = one pass of the hot path (pyramid precompute -> E2PN backbone -> superpoint transformer -> SuperPointMatching)
over a batch of `--pairs` synthetic 3DMatch-shaped pairs per GPU, random-init SE3ET-I weights.
  value : pairs/sec, inputs resident in HBM, CUDA events, max over ranks           (whole job, all GPUs)
  e2e   : pairs/sec through model.forward_pairs(host arrays): pinned H2D of the points + D2H of the correspondences
  roofline : dominant kernel (picked from the live per-entry-point event timings), achieved vs measured peak
  cpu_baseline : oracle/_ref (unmodified reference C++ precompute) + the torch-CPU port of the forward, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "3dmatch_shape_pairs_per_sec"
VARIANT = "se3eti.3dmatch"
WORKLOADS = {
    # name: (variant, metric, synthetic generator, description)
    "3dmatch": ("se3eti.3dmatch", "3dmatch_shape_pairs_per_sec", "make_3dmatch_pair",
                "~15k pts/cloud, voxel 0.025 m, 4 stages"),
    "kitti": ("se3eti.kitti", "kitti_shape_pairs_per_sec", "make_kitti_pair",
              "~30k pts/cloud, voxel 0.3 m, 5 stages"),
    # BASELINE.json configs[2]: SE3ET-E on ~5k-point pairs (a 1.5 m crop of the 3DMatch-shaped fragments)
    "3dmatch-e": ("se3ete.3dmatch", "3dmatch_5k_shape_pairs_per_sec", "make_3dmatch_pair_5k",
                  "~5k pts/cloud, voxel 0.025 m, 4 stages, SE3ET-E"),
    # BASELINE.json configs[4]: SE3ET-I training step (forward + backward), one pair per GPU per iteration as the
    # reference trains (batch_size 1), gradients averaged over the ranks by one NCCL all-reduce
    "train": ("se3eti.3dmatch", "3dmatch_shape_training_pairs_per_sec", "make_3dmatch_pair",
              "~15k pts/cloud, voxel 0.025 m, 4 stages, forward + backward + gradient all-reduce + Adam"),
}


def model_name():
    """SE3ET-I / SE3ET-E ... from the variant string (se3eti.3dmatch -> SE3ET-I)."""
    return "SE3ET-" + VARIANT.split(".")[0][len("se3et"):].upper()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([v.strip() for v in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_pairs(count, first=0, generator="make_3dmatch_pair"):
    from se3et_b200 import synthetic
    if generator == "make_3dmatch_pair_5k":
        pairs, i = [], first
        while len(pairs) < count:  # some seeds leave too little inside the crop
            try:
                p = synthetic.make_3dmatch_pair(i, crop=1.5)
                if min(len(p["ref_points"]), len(p["src_points"])) > 2000:
                    pairs.append(p)
            except ValueError:
                pass
            i += 1
    else:
        pairs = [getattr(synthetic, generator)(first + i) for i in range(count)]
    return [(p["ref_points"], p["src_points"]) for p in pairs]


def calibrated_limits(cfg, generator, dev=None):
    """KITTI neighbour limits are calibrated on the dataset by the reference (utils/data.py:212-252, keep_ratio 0.8,
    2000 samples per stage): same procedure on the synthetic KITTI-shaped pairs of seeds 0.., on the GPU path
    (se3et_b200.calibrate) or, for the CPU arm, with the reference's own C++ operators (oracle/calibrate.py); the two
    agree exactly (tests/test_points_gpu.py)."""
    b = cfg.backbone
    clouds = make_pairs(4, first=0, generator=generator)
    if dev is not None:
        from se3et_b200.calibrate import calibrate_neighbors_stack_mode
        lim = calibrate_neighbors_stack_mode(clouds, b.num_stages, b.init_voxel_size, b.init_radius, device=dev)
    else:
        from oracle import calibrate as ocal
        from oracle import points as op
        lim = ocal.calibrate_neighbors_stack_mode(clouds, b.num_stages, b.init_voxel_size, b.init_radius,
                                                  impl="ref" if op.have_ref() else "oracle")
    return [int(v) for v in lim]


NCU_KERNEL = {"se3et_kpconv_rows": "rows::kpconv_rows_kernel", "se3et_kpconv_fused": "kpconv_fused_kernel",
              "se3et_kpconv_lift": "lift::kpconv_lift_kernel", "se3et_gemm_bf16_gnstats": "gemm_tma_kernel",
              "se3et_gemm_bf16_gnapply": "gemm_stream_gnapply_kernel<0>", "se3et_gemm_grouped_bf16": "gemm_tma_kernel<32, 0>",
              "se3et_gemm_bf16_gnapply_dual": "gemm_stream_gnapply_kernel<1>", "se3et_linear_gnstats_gram": "gram_kernel",
              "se3et_linear_gnstats_stream": "gst::gnstats_stream_t_kernel",
              "se3et_geo_embed_project": "geo_embed_project_kernel", "se3et_geo_embed_lookup": "geo_embed_lookup_kernel",
              "se3et_radius_neighbors": "radius_cell_kernel", "se3et_maxpool_nbr": "maxpool_nbr_kernel",
              "se3et_groupnorm_double": "groupnorm_double_kernel", "se3et_flash_attention": "flash_attention_kernel"}


def ncu_traffic(entry_point, summary="profiles/r2_final_ncu_full_summary.csv"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes, mean over the launches of the entry point's
    kernel) from the committed `ncu --set full` capture of the same launch sequence; None when there is no capture."""
    import csv
    path = os.path.join(ROOT, summary)
    name = NCU_KERNEL.get(entry_point)
    if not name or not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    try:
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    except ValueError:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
            for r in rows[2:] if r and r[0].startswith(name)]
    return sum(vals) / len(vals) if vals else None


def algorithmic_work(cfg, n_levels, n_ref_c, n_src_c, limits):
    """Algorithmic work of one pair per entry point: ('hbm', bytes) or ('tensor', flops), from the
    layer table of the E2PN backbone (se3et_b200/modules/e2pn.py:E2PN) and the transformer shapes.
    n_levels = stacked point counts per pyramid level, limits = neighbour columns per level."""
    d = cfg.backbone.init_dim
    S = cfg.backbone.num_stages
    conv_flops = {"se3et_kpconv_rows": 0.0, "se3et_kpconv_fused": 0.0}   # tcgen05 contraction only
    norm_bytes = 0.0
    # Linear + GroupNorm passes (ops/gemm.py): statistics (GEMM epilogue or Gram matrix of the input), apply, dual apply
    ub = {"stream": 0.0, "gram": 0.0, "apply": 0.0, "dual": 0.0}

    def stats_pass(rows, k, n):
        # ops/gemm.py:linear_gn_stats: Gram matrix of the input for widening Linears with a small K, else the streaming pass
        ub["gram" if (k in (32, 64, 128) and n >= 2 * k) else "stream"] += rows * k * 2

    def block(nq, ns, h, cin, cout, strided):
        nonlocal norm_bytes
        mid = cout // 4
        if cin != mid:
            stats_pass(6.0 * ns, cin, mid)
            ub["apply"] += 6.0 * ns * (cin + mid) * 2
        # class-pre-summed contraction as issued on tcgen05 (rows = points for mid <= 128: e2pn.py:_rows_ok); the
        # mma.sync basis weighting that feeds it is NOT counted against the tcgen05 peak
        conv_flops["se3et_kpconv_rows" if mid <= 128 else "se3et_kpconv_fused"] += 2.0 * nq * 6 * 36 * mid * mid
        norm_bytes += nq * 6 * mid * (2 + 2 + 2 + 2)  # double GroupNorm: two statistics passes + apply over the bf16 conv output, bf16 out
        rows = 6.0 * nq
        stats_pass(rows, mid, cout)
        if cin != cout:
            stats_pass(rows, cin, cout)
            ub["dual"] += rows * (mid + cin + cout) * 2          # both inputs in, the block output out
        else:
            ub["apply"] += rows * (mid + 2 * cout) * 2           # input, residual in, output out

    width = 2 * d
    block(n_levels[0], n_levels[0], limits[0], d, 2 * d, False)
    for st in range(2, S + 1):
        l = st - 1
        block(n_levels[l], n_levels[l - 1], limits[l - 1], width, width, True)
        block(n_levels[l], n_levels[l], limits[l], width, 2 * width, False)
        block(n_levels[l], n_levels[l], limits[l], 2 * width, 2 * width, False)
        width *= 2
    c = cfg.geotransformer.hidden_dim
    nn2 = float(n_ref_c ** 2 + n_src_c ** 2)
    nself = sum(1 for b in cfg.geotransformer.blocks if b == "self_eq")
    # index bytes the timed path produces (precompute.py, backbone_only=True): neighbours and subsampling at full
    # width, one column for upsampling[1:], upsampling[0] skipped
    search_bytes = sum(8.0 * n_levels[i] * limits[i] for i in range(S))
    search_bytes += sum(8.0 * n_levels[i + 1] * limits[i] for i in range(S - 1))
    search_bytes += sum(8.0 * n_levels[i] for i in range(1, S - 1))
    return {
        "se3et_kpconv_rows": ("tensor", conv_flops["se3et_kpconv_rows"]),
        "se3et_kpconv_fused": ("tensor", conv_flops["se3et_kpconv_fused"]),
        "se3et_linear_gnstats_stream": ("hbm", ub["stream"]), "se3et_linear_gnstats_gram": ("hbm", ub["gram"]),
        "se3et_gemm_bf16_gnapply": ("hbm", ub["apply"]), "se3et_gemm_bf16_gnapply_dual": ("hbm", ub["dual"]),
        "se3et_groupnorm_double": ("hbm", norm_bytes),
        "se3et_geo_embed_project": ("tensor", 8.0 * nn2 * c * c),
        # tabulated embedding: index row in, bf16 row out (the table reads hit L2)
        "se3et_geo_embed_lookup": ("hbm", nn2 * (16.0 + 2.0 * c)),
        "se3et_radius_neighbors": ("hbm", search_bytes + 24.0 * sum(n_levels)),
        # positional score term: HBM-bound, every self_eq layer streams the (sum n^2, C) bf16 embedding once
        "se3et_gemm_grouped_bf16": ("hbm", nself * nn2 * (2.0 * c + 4.0 * 6 * cfg.geotransformer.num_heads)),
        "se3et_flash_attention": ("tensor", nself * 4.0 * nn2 * c * 6),
    }


# ------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: reference C++ precompute (oracle/_ref) + torch-CPU port of the forward
# ------------------------------------------------------------------------------------------------------------------
def cpu_forward_factory():
    import torch
    from oracle import e2pn as oe
    from oracle import points as op
    from oracle import transformer as ot
    from se3et_b200.model import create_model, make_cfg
    cfg = make_cfg(VARIANT)
    if VARIANT.endswith(".kitti"):
        cfg.neighbor_limits = calibrated_limits(cfg, "make_kitti_pair")
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in create_model(cfg).state_dict().items()}
    impl = "ref_raw" if op.have_ref() else "oracle"
    b, g = cfg.backbone, cfg.geotransformer

    def run(ref, src):
        pts = np.concatenate([ref, src])
        lens = np.array([len(ref), len(src)])
        d = op.precompute_data_stack_mode(pts, lens, b.num_stages, b.init_voxel_size, b.init_radius,
                                          cfg.neighbor_limits, impl=impl)
        with torch.no_grad():
            fl = oe.e2pn_forward(sd, torch.ones(len(pts), 1), d, b.init_sigma, b.group_norm)
            n = int(d["lengths"][-1][0])
            pc = torch.from_numpy(d["points"][-1])
            r, s, _, _ = ot.geometric_transformer(sd, pc[:n], pc[n:], fl[-1][:n], fl[-1][n:], g.blocks, g.hidden_dim,
                                                  g.num_heads, g.sigma_d, g.sigma_a, g.angle_k)
            r = torch.nn.functional.normalize(r, p=2, dim=1)
            s = torch.nn.functional.normalize(s, p=2, dim=1)
            return ot.superpoint_matching(r, s, torch.ones(len(r), dtype=torch.bool), torch.ones(len(s), dtype=torch.bool),
                                          cfg.coarse_matching.num_correspondences)
    return run, ("reference-c++ precompute + port forward" if impl == "ref_raw" else "port")


def run_cpu(steps, warmup, pairs_per_step=1, generator="make_3dmatch_pair"):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    run, kind_note = cpu_forward_factory()
    clouds = make_pairs(max(1, pairs_per_step), generator=generator)
    for _ in range(warmup):
        run(*clouds[0])
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        for c in clouds[:pairs_per_step]:
            run(*c)
            done += 1
    dt = time.perf_counter() - t0
    return done / dt, dt, cores, kind_note, done


def reference_arm(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), min(args.warmup, 1)
    steps = min(steps, 4)  # each step is one full pair through the CPU path (several seconds)
    value, dt, cores, note, done = run_cpu(steps, warmup, generator=WORKLOADS[args.workload][2])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s %s-shaped inference, 1 synthetic pair per step (bounded sample of the "
                               "%d-pair batch; %s), random-init weights" % (model_name(), args.workload, args.pairs,
                                                                           WORKLOADS[args.workload][3]),
                   "variant": VARIANT},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d synthetic %s-shaped pair(s): %s, torch threads = %d" % (done, args.workload, note, cores)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_extra(name, args, world, rank, dev, steps=3, warmup=2):
    """Secondary workloads of BASELINE.json (configs[2] SE3ET-E on ~5k-point pairs, configs[3] KITTI-shaped pairs), run in
    the same job at the same number of GPUs so that the driver's BENCH / SCALE records carry them: device-timed
    pairs/s (max over ranks), end to end through forward_pairs, and the dominant entry point of one serial step."""
    import torch
    import torch.distributed as dist
    from se3et_b200 import _lib, sharding
    from se3et_b200.model import create_model, make_cfg
    variant, metric, generator, note = WORKLOADS[name]
    pairs_per_gpu, ppl = (16, 8) if name == "kitti" else (64, 32)
    cfg = make_cfg(variant)
    if name == "kitti":
        cfg.neighbor_limits = calibrated_limits(cfg, generator, dev)
    torch.manual_seed(0)
    model = create_model(cfg).to(dev).eval()
    distinct = make_pairs(min(pairs_per_gpu, 16), first=0, generator=generator)
    mine = sharding.pairs_for_rank(world * pairs_per_gpu, rank, world)
    clouds = [distinct[i % len(distinct)] for i in mine]
    groups = [clouds[i:i + ppl] for i in range(0, len(clouds), ppl)]
    dev_inputs = []
    for g in groups:
        lens = np.array([len(c) for pair in g for c in pair], dtype=np.int64)
        dev_inputs.append((torch.from_numpy(np.concatenate([c for pair in g for c in pair])).to(dev), torch.from_numpy(lens)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxed(v):
        if world > 1:
            t = torch.tensor([v], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    for _ in range(warmup):
        model.forward_stacked_concurrent(dev_inputs, num_streams=args.streams)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        model.forward_stacked_concurrent(dev_inputs, num_streams=args.streams)
    e1.record()
    barrier()
    ms = maxed(e0.elapsed_time(e1))
    model.forward_pairs_concurrent(groups, num_streams=args.streams)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        model.forward_pairs_concurrent(groups, num_streams=args.streams)
    torch.cuda.synchronize()
    e2e_s = maxed(time.perf_counter() - t0)
    # one serial step with per-entry-point events
    L = _lib.lib()
    names = [n for n in _lib.KERNELS_PER_CALL]
    L.enabled = True
    L.reset(timed=names)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for pts_i, lens_i in dev_inputs:
        model.forward_stacked(pts_i, lens_i)
    s1.record()
    torch.cuda.synchronize()
    L.enabled = False
    per = {n: L.timed_ms(n)[0] for n in names}
    top = sorted(per.items(), key=lambda kv: -kv[1])[:4]
    total = world * pairs_per_gpu * steps
    return {"metric": metric, "value": total / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms / steps,
            "e2e": {"value": total / e2e_s, "unit": "pairs/s"},
            "config": {"workload": "%s %s-shaped inference (%s)" % ("SE3ET-" + variant.split(".")[0][5:].upper(), name, note),
                       "variant": variant, "pairs_per_gpu_per_step": pairs_per_gpu, "pairs_per_launch": ppl,
                       "neighbor_limits": list(cfg.neighbor_limits), "steps": steps, "warmup": warmup},
            "top_entry_points_ms": {k: round(v, 3) for k, v in top}, "serial_step_ms": round(s0.elapsed_time(s1), 3)}


def measure_train(args, world, rank, dev, steps=None, warmup=None):
    """configs[4]: training iterations of SE3ET-I on synthetic 3DMatch-shaped pairs, one pair per rank per iteration
    (se3et_b200/training.py: CUDA forward, ATen-recompute backward, reference losses, one flattened gradient all-reduce
    over NCCL, Adam).  pairs/s = world pairs per iteration / iteration time (device events, max over ranks)."""
    import torch
    import torch.distributed as dist
    from se3et_b200 import training as TR
    from se3et_b200.model import create_model, make_cfg
    steps = steps or max(2, min(args.steps, 5))
    warmup = warmup if warmup is not None else 4   # every one of the four synthetic pairs once: allocator / cuBLAS warm for its shapes
    cfg = make_cfg("se3eti.3dmatch")
    TR.RECOMPUTE['autocast'] = torch.bfloat16   # configs[4]: bf16 (operands; statistics, losses and the optimizer stay fp32)
    torch.manual_seed(0)
    model = create_model(cfg).to(dev).train()
    params = TR.trainable_parameters(model)
    opt = torch.optim.Adam(params, lr=1e-4)
    from se3et_b200 import synthetic
    mine = [synthetic.make_3dmatch_pair(100 + rank * 16 + i) for i in range(4)]
    rng = np.random.default_rng(rank)

    def step(i):
        p = mine[i % len(mine)]
        return TR.training_step(model, p["ref_points"], p["src_points"], p["transform"], optimizer=opt, world_size=world,
                                rng=rng)
    for i in range(warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(steps):
        last = step(warmup + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    nparam = sum(p.numel() for p in params)
    return {"metric": WORKLOADS["train"][1], "value": world * steps / (ms / 1e3), "unit": "pairs/s",
            "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "config": {"workload": "SE3ET-I 3dmatch-shaped training step (%s)" % WORKLOADS["train"][3],
                       "variant": "se3eti.3dmatch", "pairs_per_gpu_per_step": 1,
                       "parallelism": "data parallel over %d GPU(s), one flattened gradient all-reduce per step" % world,
                       "backward": "recompute through ATen under bf16 autocast (se3et_b200/training.py); forward on the CUDA path"},
            "grad_allreduce_bytes_per_step": int(last["grad_bytes"]) if last else 0, "parameters": int(nparam),
            "last_loss": last["loss"] if last else None}


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="pairs per step per GPU")
    ap.add_argument("--pairs-per-launch", type=int, default=32, help="pairs stacked into one launch sequence")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic pairs generated (cycled)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (KITTI-shaped, SE3ET-E)")
    ap.add_argument("--streams", type=int, default=2,
                    help="launch sequences in flight per GPU (host thread + CUDA stream each)")
    ap.add_argument("--workload", default="3dmatch", choices=sorted(WORKLOADS),
                    help="3dmatch = BASELINE.json configs[1] (the headline); kitti = configs[3]")
    args = ap.parse_args()
    global VARIANT, METRIC
    VARIANT, METRIC, generator, shape_note = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from se3et_b200 import _lib
    from se3et_b200.model import create_model, make_cfg

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)

    cfg = make_cfg(VARIANT)
    if args.workload == "kitti":
        cfg.neighbor_limits = calibrated_limits(cfg, generator, dev)
    torch.manual_seed(0)
    model = create_model(cfg).to(dev).eval()

    # the job is world * pairs synthetic pairs per step, sharded round-robin by pair (pair i -> rank i mod world, no
    # collective on the data path); `distinct` different pairs are generated and cycled
    from se3et_b200 import sharding
    distinct = make_pairs(min(args.distinct, args.pairs), first=0, generator=generator)
    mine = sharding.pairs_for_rank(world * args.pairs, rank, world)
    clouds = [distinct[i % len(distinct)] for i in mine]
    ppl = max(1, min(args.pairs_per_launch, args.pairs))
    groups = [clouds[i:i + ppl] for i in range(0, len(clouds), ppl)]
    dev_inputs = []
    for g in groups:
        lens = np.array([len(c) for pair in g for c in pair], dtype=np.int64)
        pts = torch.from_numpy(np.concatenate([c for pair in g for c in pair])).to(dev)
        dev_inputs.append((pts, torch.from_numpy(lens)))
    in_bytes = sum(int(p.numel()) * 4 for p, _ in dev_inputs)
    pinned = [torch.empty((max(int(l.sum()) for _, l in dev_inputs), 3), dtype=torch.float32).pin_memory()
              for _ in range(max(1, args.streams))]

    def step_device():
        model.forward_stacked_concurrent(dev_inputs, num_streams=args.streams)

    def step_e2e():
        out_bytes = 0
        for res in model.forward_pairs_concurrent(groups, num_streams=args.streams, pinned=pinned):
            out_bytes += sum(r[0].nbytes + r[1].nbytes + r[2].nbytes for r in res)
        return out_bytes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = _lib.lib()
    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed region: K steps, CUDA events, per-entry-point events for the roofline
    timed_names = ["se3et_kpconv_rows", "se3et_kpconv_fused", "se3et_kpconv_cin1", "se3et_kpconv_lift", "se3et_kpconv_gather", "se3et_gemm_bf16", "se3et_gemm_bf16_gnstats",
                   "se3et_gemm_bf16_gnapply", "se3et_gemm_bf16_gnapply_dual", "se3et_linear_gnstats_gram", "se3et_linear_gnstats_gram2",
                   "se3et_linear_gnstats_stream", "se3et_gemm_grouped_bf16", "se3et_groupnorm_double",
                   "se3et_groupnorm_apply", "se3et_groupnorm_stats", "se3et_maxpool_nbr", "se3et_radius_neighbors", "se3et_grid_subsample",
                   "se3et_geo_embed_project", "se3et_geo_embed_lookup", "se3et_geo_embed_indices", "se3et_flash_attention",
                   "se3et_linear_add_layernorm", "se3et_add_layernorm",
                   "se3et_superpoint_matching", "se3et_point_to_node_partition"]
    L.enabled = True
    L.reset(timed=timed_names if args.streams <= 1 else ())
    # rank 0 alone samples (all GPUs of the job, one nvidia-smi process): eight 10 Hz pollers, one per rank, hold the
    # driver's locks often enough to show up in the launch path at 8 GPUs
    sampler = ClockSampler(",".join(str(i) for i in range(world))) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    L.enabled = False
    ms = e0.elapsed_time(e1)
    launches = L.launches()
    timing_pass = "timed region"
    pass_ms, pass_pairs = ms, args.pairs * args.steps
    if args.streams > 1:
        # per-kernel durations are only meaningful without a second sequence sharing the SMs: one more step, serial,
        # with CUDA events around every entry point (same inputs, not part of `value`)
        timing_pass = "one serial step after the timed region (streams in flight make per-kernel events overlap)"
        for pts_i, lens_i in dev_inputs:  # the default stream's allocator pool is cold after the concurrent steps
            model.forward_stacked(pts_i, lens_i)
        torch.cuda.synchronize()
        L.enabled = True
        L.reset(timed=timed_names)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for pts_i, lens_i in dev_inputs:
            model.forward_stacked(pts_i, lens_i)
        s1.record()
        torch.cuda.synchronize()
        L.enabled = False
        pass_ms, pass_pairs = s0.elapsed_time(s1), args.pairs
    per_api = {n: L.timed_ms(n) for n in timed_names}
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * args.pairs * args.steps / (ms / 1e3)

    # ---- end to end through the public API (host buffers in, host results out)
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    out_bytes = 0
    for _ in range(args.steps):
        out_bytes = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * args.pairs * args.steps / e2e_s

    extra = {}
    if args.workload == "3dmatch" and not args.no_extra:
        del model, dev_inputs, pinned
        torch.cuda.empty_cache()
        for name in ("kitti", "3dmatch-e", "train"):
            try:
                extra[name] = (measure_train(args, world, rank, dev) if name == "train"
                               else measure_extra(name, args, world, rank, dev))
            except Exception as e:  # the headline line must survive a failing secondary workload
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
        model = create_model(cfg).to(dev).eval()
        dev_inputs = [(torch.from_numpy(np.concatenate([c for pair in groups[0] for c in pair])).to(dev),
                       torch.from_numpy(np.array([len(c) for pair in groups[0] for c in pair], dtype=np.int64)))]

    if rank == 0:
        peaks = load_peaks()
        # dominant entry point by summed device time inside the timed region
        pair_units = pass_pairs  # pairs this rank processed in the pass the entry-point events cover
        # algorithmic work per pair from the layer table and the measured pyramid
        pts0, lens0 = dev_inputs[0]
        dd = model.forward_stacked(pts0, lens0)["data_dict"]
        npair = lens0.shape[0] // 2
        n_levels = [float(p.shape[0]) / npair for p in dd["points"]]
        lc = dd["lengths"][-1].cpu().numpy()
        ALG = algorithmic_work(cfg, n_levels, float(lc[0::2].mean()), float(lc[1::2].mean()), cfg.neighbor_limits)
        timed = {n: v for n, v in per_api.items() if n in ALG}
        dom = max(timed, key=lambda n: timed[n][0])
        dom_ms, dom_calls = per_api[dom]
        bound, per_pair = ALG[dom]
        if bound == "hbm":
            achieved = per_pair * pair_units / (dom_ms / 1e3) / 1e9
            peak, unit = peaks["hbm_gbs"], "GB/s"
        else:
            achieved = per_pair * pair_units / (dom_ms / 1e3) / 1e12
            peak, unit = peaks["bf16_tflops"], "TFLOP/s"
        roofline = {"kernel": dom, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "traffic": ncu_traffic(dom), "peak_source": peaks["source"],
                    "share_of_step": dom_ms / pass_ms, "calls": dom_calls, "event_timing": timing_pass,
                    "per_entry_point_ms": {n: round(v[0], 3) for n, v in per_api.items()},
                    # achieved / measured peak of every entry point with a work model (same formula as `frac`)
                    "per_entry_point_frac": {
                        n: round(ALG[n][1] * pair_units / (per_api[n][0] / 1e3) /
                                 ((peaks["hbm_gbs"] * 1e9) if ALG[n][0] == "hbm" else (peaks["bf16_tflops"] * 1e12)), 4)
                        for n in timed if per_api[n][0] > 0}}
        cpu = None
        if not args.no_cpu_baseline:
            v, dt, cores, note, done = run_cpu(2, 1, generator=generator)
            cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": "%d synthetic pairs of the same workload (%s), torch threads = %d, %.1f s"
                             % (done, note, cores, dt)}
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "%s %s-shaped inference, batch of %d synthetic pairs per GPU per step "
                                   "(%s), random-init weights" % (model_name(), args.workload, args.pairs, shape_note),
                       "variant": VARIANT, "pairs_per_gpu_per_step": args.pairs, "pairs_per_launch": ppl,
                       "distinct_pairs": len(distinct), "neighbor_limits": list(cfg.neighbor_limits),
                       "streams_per_gpu": args.streams, "parallelism": "pairs sharded over %d GPU(s), no collective" % world,
                       "l2": "working set per launch (activations of %d stacked pairs, > 1 GB) exceeds the 126 MB L2" % ppl},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes},
            "roofline": roofline, "cpu_baseline": cpu, "extra_workloads": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
