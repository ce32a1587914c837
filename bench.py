#!/usr/bin/env python
"""bench.py -- matched image pairs / second of the Matcher hot path (coarse match -> window gather -> fine match).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path (one JSON line)
    python bench.py --impl reference [...]                        # the reference's CPU path (oracle port), rank 0 only

Workload (BASELINE.json configs[1]): a batch of 64 synthetic 480x640 pairs per GPU per step -- coarse features
[64, 4800, 256] x2 and fine maps [64, 128, 240, 320] x2 (channels-last), bf16, planted correspondences
(pope_b200/synth.py).  A "step" is one pass of the hot path over that batch.

  value : pairs/s with the inputs already resident in HBM (CUDA events, max over ranks).  The batch's inputs
          (2.8 GB) are far larger than the 126 MB L2, so every step streams them from HBM.  Consecutive steps go through
          pope_b200.driver.DeviceBatchRunner: they alternate between two CUDA streams (own scratch each), so the small latency-bound kernels that end a step (column-sum
          reduction, list evaluation, compaction) overlap the next step's sweep; stage_ms / roofline come from a separate
          pass on one stream, where events bracket the kernels and not the queue.
  e2e   : the same metric through the C-ABI host entry (pope_pipeline_run): pinned host buffers in, pinned host
          buffers out, host<->device copies inside the timed region.
  roofline : the coarse stage (dominant) against the measured bf16 tensor peak; algorithmic work 2*L*S*C per pair.
             roofline_fine: the fused window-gather + fine-match kernel against the measured HBM copy bandwidth.
  cpu_baseline : the oracle port (same op sequence as the reference, torch CPU) on a bounded sample, rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "matched_pairs_per_sec_480x640"
UNIT = "pairs/s"
H, W_IMG = 480, 640
HC, WC = H // 8, W_IMG // 8          # 60 x 80 coarse cells
L = HC * WC
C_COARSE, C_FINE, FINE_STRIDE, WIN = 256, 128, 4, 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="pairs per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--coarse-impl", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-sample-pairs", type=int, default=2)
    ap.add_argument("--in-matcher", type=int, default=0, metavar="PAIRS",
                    help="also time steps 3-5 of Matcher.forward on PAIRS pairs with the PyTorch FinePreprocess Linears and "
                         "fine transformer between the CUDA stages (fp32, SURVEY 8(d) 'in-Matcher' figure)")
    ap.add_argument("--streams", type=int, default=2,
                    help="device-resident steps alternate between this many CUDA streams (each with its own scratch), so "
                         "that the small latency-bound kernels at the end of a step overlap the next step's sweep")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled in the background from before the warm-up until the end of the
    run; `summary(t0, t1)` reports the samples that fall inside the timed region (wall-clock window)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, period_ms: int = 20):
        import threading
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.samples.append((time.time(), ln))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc is not None and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ts, ln in list(self.samples):
            f = [x.strip() for x in ln.split(",")]
            try:
                rows.append((ts, float(f[0]), float(f[1]), float(f[2]), [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        inside = [r for r in rows if t0 - 0.03 <= r[0] <= t1 + 0.03]
        scope = "timed region"
        if not inside:       # region shorter than one sampling period: fall back to the samples taken under load
            pmax = max(r[3] for r in rows)
            inside = [r for r in rows if r[3] >= 0.6 * pmax]
            scope = "samples under load (timed region shorter than the sampling period)"
        reasons = sorted({x for r in inside for x in r[4]})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": max(r[2] for r in rows),
                "power_w_max": max(r[3] for r in inside), "reasons": reasons, "samples": len(inside), "scope": scope}


def cpu_reference_pairs_per_sec(n_sample: int, repeats: int = 1):
    """The reference's CPU path for the hot path (oracle port: einsum -> softmax x softmax -> ... -> unfold -> gather ->
    fine match), fp32, all host threads, on `n_sample` pairs of the bench workload."""
    from oracle import pope_oracle as O
    from pope_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    f0, f1 = synth.coarse_features(1234, n_sample, L, L, C_COARSE)
    ff0, ff1 = synth.fine_feature_maps(4321, n_sample, HC * FINE_STRIDE, WC * FINE_STRIDE, C_FINE, channels_last=False)
    best, m = float("inf"), 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = O.match_pairs(f0, f1, ff0, ff1, (H, W_IMG), (HC, WC), (HC, WC))
        best = min(best, time.perf_counter() - t0)
        m = out["b_ids"].numel()
    return n_sample / best, best, m, torch.get_num_threads()


def in_matcher_figure(n_pairs: int, dev):
    """Steps 3-5 of Matcher.forward (src/matcher/matcher.py:71-79) with the drop-in modules, i.e. WITH the FinePreprocess
    Linears and the fine transformer between coarse and fine matching (SURVEY 8(d) 'in-Matcher' figure), two ways:
      fp32      : CUDA coarse (fp32-FMA path) -> CUDA gather -> torch Linears -> torch fine transformer -> CUDA fine match
      bf16_cuda : Matcher(config, fine_cuda_bf16=True) on bf16 features: tcgen05 coarse -> bf16 gather -> CUDA Linears ->
                  CUDA fine transformer (csrc/fine_tf.cu) -> CUDA fine match"""
    import pope_b200
    from pope_b200 import synth
    hf, wf = HC * FINE_STRIDE, WC * FINE_STRIDE
    shapes = {"hw0_i": torch.Size([H, W_IMG]), "hw1_i": torch.Size([H, W_IMG]), "hw0_c": torch.Size([HC, WC]),
              "hw1_c": torch.Size([HC, WC]), "hw0_f": torch.Size([hf, wf]), "hw1_f": torch.Size([hf, wf]), "bs": n_pairs}
    out = {}
    for name, dtype, reps in (("fp32", torch.float32, 2), ("bf16_cuda", torch.bfloat16, 5)):
        torch.manual_seed(0)
        m = pope_b200.Matcher(pope_b200.make_default_cfg(), fine_cuda_bf16=(dtype == torch.bfloat16)).eval().to(dev)
        f0, f1 = synth.coarse_features(99, n_pairs, L, L, C_COARSE, dtype=dtype)
        g = torch.Generator(device=dev).manual_seed(98)
        ff0 = torch.randn(n_pairs, hf, wf, C_FINE, device=dev, generator=g).to(dtype).permute(0, 3, 1, 2)
        ff1 = torch.randn(n_pairs, hf, wf, C_FINE, device=dev, generator=g).to(dtype).permute(0, 3, 1, 2)
        f0, f1 = f0.to(dev), f1.to(dev)

        def run():
            data = dict(shapes)
            with torch.no_grad():
                m.coarse_matching(f0, f1, data)
                w0, w1 = m.fine_preprocess(ff0, ff1, f0, f1, data)
                if w0.size(0):
                    w0, w1 = m.loftr_fine(w0, w1)
                m.fine_matching(w0, w1, data)
            return data

        for _ in range(2):
            data = run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            data = run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"value": n_pairs / (ms / 1e3), "unit": UNIT, "pairs": n_pairs, "ms": ms,
                     "matches": int(data["mconf"].numel())}
        del m, f0, f1, ff0, ff1, data
        torch.cuda.empty_cache()
    out["note"] = ("steps 3-5 of Matcher.forward incl. FinePreprocess Linears + fine transformer; the module flow has one host "
                   "sync per call (the match count), like the reference's torch.where")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, m = [], 0
    n = args.cpu_sample_pairs
    for it in range(args.warmup + args.steps):
        _, dt, m, cores = cpu_reference_pairs_per_sec(n)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    v = n / (ms / 1e3)
    sample = f"{n} pairs/step of the 480x640 workload (fp32, torch CPU ops of the reference path, M={m} matches)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "480x640 pairs, coarse 60x80 tokens d=256 + fine 5x5 windows d=128 (BASELINE configs[1])",
                   "pairs_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from pope_b200 import _lib, driver, ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path to fall back to)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    impl = {"auto": _lib.COARSE_AUTO, "simt": _lib.COARSE_SIMT, "tcgen05": _lib.COARSE_TCGEN05}[args.coarse_impl]
    n = args.pairs
    esize = 2 if dtype == torch.bfloat16 else 4

    # ---- synthetic inputs (seeded per rank: every rank owns different pairs) --------------------------------------
    f0, f1 = synth.coarse_features(1234 + rank, n, L, L, C_COARSE, dtype=dtype)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    hf, wf = HC * FINE_STRIDE, WC * FINE_STRIDE
    ff0 = torch.randn(n, hf, wf, C_FINE, device=dev, generator=g).to(dtype).permute(0, 3, 1, 2)   # channels-last
    ff1 = torch.randn(n, hf, wf, C_FINE, device=dev, generator=g).to(dtype).permute(0, 3, 1, 2)
    d_f0, d_f1 = f0.to(dev), f1.to(dev)
    # N > 1: every step appends its packed match records to a device buffer; the job's ONE cross-GPU step (a single
    # gather of the live records) runs at the end of the K timed steps, inside the timed region
    job = driver.JobGather(max(args.steps, args.warmup, 3), n * L, dev) if world > 1 else None
    n_streams = max(1, args.streams)
    wss = [torch.empty(_lib.lib().pope_coarse_workspace_bytes_ex(n, L, L, C_COARSE, _lib.dtype_code(d_f0)), dtype=torch.uint8,
                       device=dev) for _ in range(n_streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)] if n_streams > 1 else [torch.cuda.current_stream(dev)]
    step_no = [0]

    last_add = [None]                      # event after the previous step's append to the job's record buffer
    runner = driver.DeviceBatchRunner(dev, n_streams)      # the public form of "consecutive batches on alternating streams"

    def add_to_job(res):                   # appends happen in step order: chain them with events across the streams
        st = torch.cuda.current_stream(dev)
        if last_add[0] is not None:
            st.wait_event(last_add[0])
        job.add(res, rank * n)
        last_add[0] = torch.cuda.Event()
        last_add[0].record(st)

    def step_device(ev=None):
        if ev is not None:                 # stage-timing pass: one stream, events around the two stages
            with torch.cuda.stream(streams[0]):
                return _step_on_stream(ev, wss[0])
        res, _ = runner.submit(d_f0, d_f1, ff0, ff1, (H, W_IMG), (HC, WC), (HC, WC), impl=impl,
                               after=add_to_job if job is not None else None)
        return res

    def join_streams():
        runner.join()

    def fork_streams():
        runner.fork()

    def _step_on_stream(ev, ws):
        if ev: ev[0].record()
        res = ops.coarse_match(d_f0, d_f1, (HC, WC), (HC, WC), 8.0, impl=impl, workspace=ws)
        if ev: ev[1].record()
        m_dev = res["counts"][n:n + 1]
        # fused window gather + fine match (the hot-path-only pipeline has no fine transformer in between)
        expec, mk1f = ops.fine_match_maps(ff0, ff1, res["b_ids"], res["i_ids"], res["j_ids"], res["mkpts1_c"], WC, WC,
                                          FINE_STRIDE, (WIN // 2) * 2.0, WIN, m_dev)
        if ev: ev[2].record()
        res.update(mkpts1_f=mk1f, mkpts0_f=res["mkpts0_c"], expec_f=expec)
        return res


    def gather_step(res):
        pass                               # (the append to the job's record buffer is part of step_device)

    def finish_job():
        join_streams()
        if job is not None:
            job.finish(rank, world)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    clk = ClockSampler(local)
    clk.wait_first()
    for _ in range(max(args.warmup, 3)):
        gather_step(step_device())
    finish_job()
    barrier()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_beg.record()
    fork_streams()
    for k in range(args.steps):
        res = step_device()
        gather_step(res)
    finish_job()
    t_end.record()
    barrier()
    wall1 = time.time()
    total_ms = t_beg.elapsed_time(t_end)
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n * args.steps / (total_ms / 1e3)
    # per-stage durations for the rooflines: a few extra steps on ONE stream (with several streams the stages of
    # consecutive steps overlap, so events around them would time the queue, not the kernels); not part of `value`
    k_stage = max(4, min(args.steps, 6))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(k_stage)]
    for k in range(k_stage):
        res = step_device(evs[k])
    torch.cuda.synchronize(dev)
    if os.environ.get("POPE_BENCH_DEBUG"):
        print("stage pass:", [(round(e[0].elapsed_time(e[1]), 3), round(e[1].elapsed_time(e[2]), 3)) for e in evs], file=sys.stderr)
    # the first step of the pass starts on an idle stream: its first event also times the host's launch latency
    coarse_ms = statistics.median(e[0].elapsed_time(e[1]) for e in evs[1:])
    fine_ms = statistics.median(e[1].elapsed_time(e[2]) for e in evs[1:])
    M = res.total()
    flags = res.flags()

    # ---- end to end through the C-ABI host entry (pinned host buffers, copies inside the timed region) -------------
    e2e = None
    if not args.no_e2e:
        chunk = min(16, n)
        pl = driver.Pipeline(dtype, chunk, (H, W_IMG), (HC, WC), (HC, WC), C_COARSE, C_FINE, FINE_STRIDE, WIN, impl=impl,
                             device=local)
        prev_aff = driver.bind_host_to_gpu(local)        # page-locked staging buffers on the GPU's NUMA node
        h_f0, h_f1 = f0.pin_memory(), f1.pin_memory()
        h_ff0 = ff0.permute(0, 2, 3, 1).contiguous().cpu().pin_memory()
        h_ff1 = ff1.permute(0, 2, 3, 1).contiguous().cpu().pin_memory()
        out = pl.alloc_outputs(n)
        if prev_aff is not None:
            os.sched_setaffinity(0, prev_aff)
        for _ in range(2):
            pl.run(h_f0, h_f1, h_ff0, h_ff1, out)
        barrier()
        k_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            pl.run(h_f0, h_f1, h_ff0, h_ff1, out)       # returns after the results are in host memory
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = pl.last_h2d_bytes      # bulk copies + the centre pixels of image 0 read in place from pinned memory
        d2h = sum(out[k].numel() * out[k].element_size() for k in ("i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f", "counts")) + 4 * ((n + chunk - 1) // chunk)
        e2e = {"value": world * n * k_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e, "matches": int(out["counts"].sum()),
               "api": "pope_pipeline_run (C ABI, pinned host buffers, chunk=%d pairs)" % chunk}
        assert int(out["counts"].sum()) == M, (int(out["counts"].sum()), M)
        pl.close()

    clk.stop()
    clocks = clk.summary(wall0, wall1)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, tf_burst, tf_sus, peak_src = measured_peaks()
    # DRAM traffic per launch from the committed ncu --set full capture of this workload
    # (profiles/r1_ncu_single_sweep_step.csv: dram__bytes_read.sum + dram__bytes_write.sum; coarse = single sweep +
    # column-sum reduction + gated launch + list evaluation + count/emit).  The sweep streams the 315 MB of coarse
    # features once and writes 170 MB of per-32-row column partial sums that the reduction reads back.
    std_workload = n == 64 and dtype == torch.bfloat16
    coarse_traffic = 721.9e6 if std_workload else None
    fine_traffic = 850.3e6 if std_workload else None
    flops = n * 2.0 * L * L * C_COARSE
    ach = flops / (coarse_ms / 1e3) / 1e12
    fine_bytes = M * ((1 + 25) * C_FINE * esize + 3 * 8 + 8 + 12 + 8)     # 26 feature rows + ids + coords in/out
    tc = impl == _lib.COARSE_TCGEN05 or (impl == _lib.COARSE_AUTO and _lib.tcgen05_available())   # fp32: split path
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "batch of 64 pairs at 480x640, coarse+fine matching (BASELINE configs[1])" if n == 64 else
                   f"batch of {n} pairs at 480x640, coarse+fine matching",
                   "pairs_per_gpu_per_step": n, "coarse_tokens": [HC, WC], "d_coarse": C_COARSE, "d_fine": C_FINE,
                   "window": WIN, "coarse_impl": ("tcgen05" if dtype == torch.bfloat16 else "tcgen05 on a three-way bf16 split of the fp32 features") if tc else "simt-fp32fma", "fine_map_layout": "channels_last",
                   "l2_policy": "inputs_exceed_l2 (2.8 GB of features per step vs 126 MB L2)",
                   "streams": n_streams,
                   "matches_per_step": M, "flags": flags, "gather": "one NCCL all-gather of the job's live match records after the K steps (inside the timed region)" if world > 1 else "none"},
        "clocks": clocks,
        "stage_ms": {"coarse": coarse_ms, "fine_gather_match_fused": fine_ms},
        "roofline": {"kernel": "coarse stage (tcgen05 single sweep: row sums + shuffle-reduced column sums + candidate lists; column-sum reduction, list evaluation, compaction)" if tc else
                               "coarse stage (fp32-FMA sweeps + compaction)", "bound": "tensor",
                     "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s", "frac": ach / tf_sus, "traffic": coarse_traffic,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "algorithmic_flops_per_launch": flops},
        "roofline_fine": {"kernel": "fine_match_maps_kernel (fused 5x5 window gather + correlation + softmax expectation)",
                          "bound": "hbm", "achieved": fine_bytes / (fine_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                          "frac": fine_bytes / (fine_ms / 1e3) / 1e9 / hbm, "traffic": fine_traffic,
                          "algorithmic_bytes_per_launch": fine_bytes},
        "gpu_launches": args.steps * _lib.KERNELS_PER_STEP[("tcgen05" if dtype == torch.bfloat16 else "tcgen05_f32") if tc else "simt"],
    }
    if args.in_matcher > 0:
        line["in_matcher"] = in_matcher_figure(args.in_matcher, dev)
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu:
        v, dt, m_cpu, cores = cpu_reference_pairs_per_sec(args.cpu_sample_pairs)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample_pairs} pairs of the same 480x640 workload, fp32 torch CPU "
                                          f"(oracle port of the reference op sequence), {dt:.2f} s, M={m_cpu}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
