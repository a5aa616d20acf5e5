#!/usr/bin/env python
"""Benchmark of the CWT per-episode head (BASELINE.json metric: episodes/sec, PASCAL 1-shot, 60x60x512).

    python bench.py --gpus N --steps K --warmup W                 # our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # reference algorithm on the host CPU

A "step" is one pass of the head (fit 200 SGD steps -> transformer -> logits/upsample/argmax/IoU) over one
batch of E synthetic episodes per GPU. N > 1: one process per GPU under torchrun, episodes sharded across ranks
(weak scaling: E per rank fixed), the only collective is the int64 all-reduce of the IoU table each step.

One JSON line on stdout (rank 0). `value` = device-resident throughput; `e2e` = the same through the public API
with pinned host buffers (H2D of every input + D2H of the counts inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "episodes_per_sec_head_pascal_1shot_60x60x512"
UNIT = "episodes/s"


def metric_name(a):
    return METRIC if a.shot == 1 else METRIC.replace("1shot", f"{a.shot}shot")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--episodes", type=int, default=64, help="episodes per GPU per step")
    ap.add_argument("--shot", type=int, default=1)
    ap.add_argument("--heads", type=int, default=4)          # scripts/test.sh:16
    ap.add_argument("--cls-lr", type=float, default=0.1)     # scripts/test.sh:15
    ap.add_argument("--adapt-iter", type=int, default=200)   # config_files/pascal.yaml:43
    ap.add_argument("--style", default="unit", choices=["unit", "backbone"])
    ap.add_argument("--fit-algo", type=int, default=0)
    ap.add_argument("--attn-algo", type=int, default=0)
    ap.add_argument("--cpu-baseline-episodes", type=int, default=8)
    ap.add_argument("--ref-episodes", type=int, default=2, help="--impl reference: episodes per step (bounded CPU sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sub-batch", type=int, default=32, help="HostPipeline sub-batch size (0 = whole batches)")
    ap.add_argument("--e2e-sub-all", type=int, default=1, help="1: every host batch goes through in sub-batches; 0: only the first")
    ap.add_argument("--e2e-expand-main", type=int, default=2, help="zero-compressed e2e: expansion kernel on the head's stream (1), the copy stream (0) or a third stream (2)")
    ap.add_argument("--e2e-ramp", default="8,8,16", help="sub-batch sizes at the start of an e2e run (first host batch only); '' = none")
    ap.add_argument("--e2e-format", default="zc", choices=["zc", "dense"], help="host format of the e2e leg: zero-compressed "
                                                                               "features (default) or dense tensors")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic episodes generated per rank (tiled to E)")
    ap.add_argument("--attn-both", action="store_true", help="also time the tcgen05 K-projection path of the transformer block "
                                                             "(reported under roofline.transformer_tcgen05; used by the ncu capture)")
    return ap.parse_args()


GEOM = dict(C=512, h=60, w=60, H=473, W=473)


def workload_name(a):
    return (f"PASCAL-5i {a.shot}-shot head, synthetic 473x473 episodes, f[{a.shot},512,60,60] fp32, heads={a.heads}, "
            f"cls_lr={a.cls_lr}, adapt_iter={a.adapt_iter} (config_files/pascal.yaml + scripts/test.sh overrides)")


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def window(self, t0, t1):
        """only samples taken inside [t0, t1] (time.time()) count: the timed region"""
        self.t0, self.t1 = t0, t1

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= self.t1)] or [r for _, r in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_kernel_counters(kernel=None):
    """profiles/*_kernel_counters.json (written by tools/ncu_counters_to_json.py from an ncu capture of this very command)
    -> (dict, file name, whether it was captured on the build that is running: the fingerprint of the whole library, or of
    the sources ``kernel`` is compiled from, is unchanged)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernel_counters.json")))
    if not files:
        return None, None, False
    try:
        from few_shot_seg_cwt_b200 import build as B
        fp = B.fingerprint()
        loaded = [(f, json.load(open(f))) for f in files]

        def same_build(d):
            # the whole library unchanged, or at least the sources of `kernel` (its .cu and the headers it includes)
            if d.get("build_fingerprint") == fp:
                return True
            kf = d.get("kernel_source_fingerprints") or {}
            return bool(kernel) and kernel in kf and kf[kernel] == B.kernel_fingerprint(kernel)

        # the capture taken on THIS build if there is one, else the newest
        f, d = next(((f, d) for f, d in reversed(loaded) if same_build(d)), loaded[-1])
        return d, os.path.relpath(f, ROOT), same_build(d)
    except Exception:
        return None, None, False


def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_episode_runner(a):
    """The reference head for one episode on the host CPU -> (callable(ep) -> outputs, kind, description).
    kind "reference": the episode body of src/test.py:162-234 around the reference's OWN MultiHeadAttentionOne and
    batch_intersectionAndUnionGPU (oracle/_ref, the verbatim copy made by oracle/make_ref.py — it travels to the GPU box);
    kind "port": the line-by-line restatement oracle/head_ref.episode_ref, when oracle/_ref has not been made."""
    from few_shot_seg_cwt_b200 import synthetic as syn
    from oracle import head_ref as O
    from oracle import ref_episode as R
    params = syn.make_transformer_params(a.heads, 512)
    mods = R.load_reference_modules(prefer_live=False)          # never reads /root/reference at run time
    if mods is not None:
        MHA, batch_iou = mods[0], mods[1]
        return (lambda ep: R.episode_via_reference(ep, params, a.heads, a.cls_lr, a.adapt_iter, MHA, batch_iou), "reference",
                "src/test.py:162-234 episode body with the reference's own src/model/transformer.py + src/util.py (oracle/_ref)")
    return (lambda ep: O.episode_ref(ep.f_s, ep.s_label, ep.f_q, ep.q_label, ep.w0, params, a.heads, a.cls_lr, a.adapt_iter), "port",
            "oracle/head_ref.episode_ref (restatement of src/test.py:162-234)")


def cpu_reference_rate(a, n_episodes: int, warm: int = 1, first_index: int = 0):
    """The reference head on the host CPU (all host threads) on the same workload: episodes ``first_index ...`` of the
    synthetic generator, i.e. the very episodes of the GPU batch. Returns (episodes/s, seconds, the outputs of the first
    episode — what `parity_check` compares the GPU result with —, kind, description)."""
    use_all_host_threads()
    from few_shot_seg_cwt_b200 import synthetic as syn
    run, kind, desc = cpu_episode_runner(a)
    eps = [syn.make_episode(first_index + i, shot=a.shot, style=a.style, **GEOM) for i in range(warm + n_episodes)]
    outs = [run(ep) for ep in eps[:warm]]
    t0 = time.perf_counter()
    for ep in eps[warm:]:
        outs.append(run(ep))
    dt = time.perf_counter() - t0
    return n_episodes / dt, dt, outs[0], kind, desc


def run_reference(a, rank):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores (the reference's own
    modules from oracle/_ref when present, else the oracle port). Rank 0 only."""
    if rank != 0:
        return
    per_step = max(1, a.ref_episodes)
    use_all_host_threads()
    from few_shot_seg_cwt_b200 import synthetic as syn
    run_ep, kind, desc = cpu_episode_runner(a)
    eps = [syn.make_episode(i, shot=a.shot, style=a.style, **GEOM) for i in range(per_step)]
    run = lambda: [run_ep(ep) for ep in eps]
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = time.perf_counter() - t0
    v = per_step * a.steps / dt
    cores = torch.get_num_threads()
    emit(json.dumps({
        "impl": "reference", "metric": metric_name(a), "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "episodes_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{per_step} episodes/step x {a.steps} steps, torch {torch.__version__} CPU fp32, {desc}"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
_JSON_FD = None


def emit(line: str) -> None:
    """The ONE JSON line goes to the process's original stdout."""
    data = (line + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line + "\n"); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    a = parse()
    # stdout carries exactly one JSON line: anything native libraries print on fd 1 during the run (NCCL's
    # "NCCL version ..." banner when NCCL_DEBUG=VERSION is set in the environment) is sent to stderr instead
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return

    import torch.distributed as dist
    import few_shot_seg_cwt_b200 as cwt
    from few_shot_seg_cwt_b200 import synthetic as syn
    from few_shot_seg_cwt_b200 import _lib as L

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback). Use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cwt.bind_host_to_gpu(dev)        # pinned staging buffers on the GPU's NUMA node (no-op when the CPUs are not visible)
    L.load()

    E = a.episodes
    F_bytes = 3600 * 512 * 4
    P = 473 * 473
    params = {k: v.to(dev) for k, v in syn.make_transformer_params(a.heads, 512).items()}

    # synthetic episodes: rank r owns indices r, r+world, ... ; `distinct` generated, tiled up to E
    nd = min(a.distinct, E)
    idx = syn.shard_indices(nd * world, rank, world)
    host = syn.make_batch(idx, shot=a.shot, style=a.style, **GEOM)
    rep = (E + nd - 1) // nd
    tile = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:E].contiguous()
    host = syn.EpisodeBatch(*(tile(t) for t in (host.f_s, host.s_label, host.f_q, host.q_label, host.w0, host.subcls, host.idx)))
    host = host.pin_memory()
    devb = host.to(dev)
    table = cwt.IoUTable(5, dev)
    torch.cuda.synchronize()

    # cwt.HeadPipeline: the fit of step i+1 is queued on the main stream while the post stage of step i (transformer,
    # fused logits / IoU, table update and the one collective — an int64 all-reduce of the counts, no-op at world 1) runs on
    # a side stream; every step's work completes inside the timed region (finish() before the closing event)
    head = cwt.HeadPipeline(dev, params, a.heads, a.cls_lr, a.adapt_iter, fit_algo=a.fit_algo, attn_algo=a.attn_algo,
                            table=table, reduce_every_step=True)

    def step_resident():
        return head.submit(devb.f_s, devb.s_label, devb.f_q, devb.q_label, devb.w0, devb.subcls)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        step_resident()
    head.finish()
    barrier()
    n0 = L.launch_count()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        last_out, _ = step_resident()
    head.finish()
    ev1.record()
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = L.launch_count() - n0
    sampler.window(t_wall0, time.time())
    clocks = sampler.stop() if rank == 0 else None
    value = world * E * a.steps / (ms / 1e3)

    # ---- per-stage timing on the launching stream (roofline) ----
    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    fit_ms, w_fit = timed(lambda: cwt.fit_classifier(devb.f_s, devb.s_label, devb.w0, a.cls_lr, a.adapt_iter, check=False,
                                                     algo=a.fit_algo))
    tr_ms, w_ad = timed(lambda: cwt.transformer_forward(w_fit, devb.f_q, params["w_qkvs.weight"], params["fc.weight"],
                                                        params["fc.bias"], params["layer_norm.weight"],
                                                        params["layer_norm.bias"], a.heads, normalize_k=True, algo=a.attn_algo))
    wts = torch.stack([w_ad, w_fit], 1)
    iou_ms, _ = timed(lambda: cwt.logits_iou(wts, devb.f_q, devb.q_label, 0b01, return_logits=False))
    label_bytes = devb.s_label.element_size()
    T, S = a.adapt_iter, a.shot
    bytes_fit = E * ((2 * T + 1) * S * F_bytes + S * P * label_bytes + 2 * (2 * 512 * 4))     # (2T+1) S F per episode
    bytes_iou = E * (F_bytes + P * label_bytes + 2 * 6 * 8)
    # bytes_transformer(E) = E*F + nH*(2*C*C*4) + E*(2*2*C*4)  (f_q once; the re-associated path reads it twice)
    bytes_tr = E * F_bytes + a.heads * (2 * 512 * 512 * 4) + E * (2 * 2 * 512 * 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    fit_gbs = bytes_fit / (fit_ms / 1e3) / 1e9
    resident = (a.fit_algo in (0, 2) and S == 1)
    sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    fit_kernel = "k_fit_resident" if resident else ("k_fit_l2" if a.fit_algo in (0, 3) else None)
    counters, counters_file, counters_match = load_kernel_counters(fit_kernel)
    kc = (counters or {}).get("kernels", {}).get(fit_kernel or "", None)
    # DRAM traffic of the dominant kernel per launch: ncu dram__bytes_read + dram__bytes_write of the SAME command
    # (tools/ncu_bench_kernels.sh -> profiles/*_kernel_counters.json, with the fingerprint of the build it was taken on)
    traffic = (kc["dram_bytes_read"] + kc["dram_bytes_write"]) if (kc and E == 64) else None
    traffic_src = ({"file": counters_file, "build_fingerprint": counters["build_fingerprint"], "matches_this_build": counters_match}
                   if traffic is not None else None)
    if resident:
        # On-chip roof: every SGD step sweeps the resident tile twice (P1 and P3): algorithmic on-chip bytes = E * 2T * F.
        # The tile lives in TENSOR MEMORY (tcgen05.st once per episode, tcgen05.ld every sweep), so the bounding resource is the
        # tensor-memory read path: peak = the tcgen05.ld rate measured by tools/micro/tmem_bench.cu on this pool's B200s
        # (profiles/r2g_tmem_bench.txt, 16 warps, 3 loads in flight) x all SMs x the SM clock sampled in the timed region.
        # CWT_RESIDENT_TMEM=0 runs the round-1 kernel (tile in shared memory, 128 B/clk/SM).
        tmem_kernel = os.environ.get("CWT_RESIDENT_TMEM", "1") != "0"
        tmem_rate, tmem_src = 378.3, "fallback constant (profiles/r2g_tmem_bench.txt as committed)"
        try:
            for line in open(os.path.join(ROOT, "profiles", "r2g_tmem_bench.txt")):
                if line.startswith("tmem only, 16 warps, 3 ld in flight"):
                    tmem_rate = float(line.split("tmem")[2].split("B/clk")[0])
                    tmem_src = "profiles/r2g_tmem_bench.txt (tools/micro/tmem_bench.cu: tcgen05.ld 32x32b.x32, 16 warps, 3 loads in flight)"
        except Exception:
            pass
        onchip_bytes = E * 2 * T * S * F_bytes
        onchip_gbs = onchip_bytes / (fit_ms / 1e3) / 1e9
        smem_peak = 128.0 * n_sm * sm_clk * 1e6 / 1e9
        tmem_peak = tmem_rate * n_sm * sm_clk * 1e6 / 1e9
        steps_per_group = T * ((E + 3) // 4)
        clk_per_step = fit_ms * 1e-3 * sm_clk * 1e6 / max(steps_per_group, 1)
        smem_equiv = {"achieved": onchip_gbs, "peak": smem_peak, "unit": "GB/s", "frac": onchip_gbs / smem_peak,
                      "peak_source": f"128 B/clk/SM x {n_sm} SMs x {sm_clk:.0f} MHz",
                      "note": "the same bytes against the shared-memory pipe: the roof of a shared-memory-resident tile (the round-1 "
                              "kernel, frac 0.30 there); kept so that the rounds compare on one denominator"}
        roofline = {
            "kernel": (("k_fit_resident<512 compute threads + halo warp, 1 CTA/SM, C=512, tile 20x5 of 60x60, tile in TENSOR MEMORY> (one "
                        "cooperative launch: features staged once by tensor-map TMA copies and moved to tensor memory with tcgen05.st, 200 SGD steps "
                        "on chip with tcgen05.ld sweeps, all-reduce through 64-bit L2 atomics polled by the compute threads)")
                       if tmem_kernel else
                       ("k_fit_resident<512 compute threads, 1 CTA/SM, C=512, tile 20x5 of 60x60> (one cooperative launch: features "
                        "staged once into shared memory by bulk-TMA, 200 SGD steps on chip, all-reduce through 64-bit L2 atomics)")),
            "bound": "tmem" if tmem_kernel else "smem", "achieved": onchip_gbs, "peak": tmem_peak if tmem_kernel else smem_peak,
            "unit": "GB/s", "frac": onchip_gbs / (tmem_peak if tmem_kernel else smem_peak),
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": (f"{tmem_rate:.1f} B/clk/SM x {n_sm} SMs x {sm_clk:.0f} MHz; rate from {tmem_src}" if tmem_kernel
                            else smem_equiv["peak_source"]),
            "note": ("the step is two sweeps of the resident tile (P1: z = Wd.F, P3: dW = g.F^T); achieved = E*2T*F algorithmic on-chip "
                     "bytes / CUDA-event time of the fit call; the kernel occupies 144 of the SMs (4 groups x 36 CTAs). The kernel is "
                     "NOT bound by this bandwidth: per step ~4 700 clk are work (P1 is issue-bound: butterfly reduce over the lanes; "
                     "full-resolution stage; gather) and ~2 400 clk are the two cross-CTA exchanges (DESIGN.md 4.4)"),
            "clk_per_step": clk_per_step,
            "B_per_clk_per_active_SM": (2 * 512 * 100 * 4) / clk_per_step,
            "smem_equivalent": smem_equiv,
            "hbm_equivalent": {"achieved": fit_gbs, "peak": hbm_peak, "unit": "GB/s", "x_of_hbm_peak": fit_gbs / hbm_peak,
                               "peak_source": peak_src,
                               "note": "SURVEY §8d's (2T+1)*S*F bytes per episode / time: what an HBM-streamed fit would have to move — "
                                       "above 1 x the HBM peak by construction, because the bytes never leave the chip"},
            "dram_frac": (traffic / (fit_ms / 1e3) / 1e9 / hbm_peak) if traffic else None,
            "compulsory_dram_bytes": E * S * F_bytes,
            # the same figures as flat keys (a flattened view of this block keeps them): the fraction on the round-1 denominator
            # (shared-memory pipe), the multiple of the HBM-streamed floor, and the issue-slot utilisation ncu measured for this
            # kernel — the resource it is closest to (it is bound by exchange latencies, not by a bandwidth)
            "frac_smem_equivalent": smem_equiv["frac"],
            "x_of_hbm_streamed_floor": fit_gbs / hbm_peak,
            "issue_slots_busy_frac": (kc["issue_active_pct"] / 100.0) if (kc and kc.get("issue_active_pct") is not None) else None,
        }
    elif fit_kernel == "k_fit_l2":
        # L2 roof: the episodes in flight stay L2-resident for all 2T sweeps; peak = L2 read bandwidth measured live with the
        # library's own microbenchmark (cwt_debug_l2_read: every CTA sweeps a 48 MB buffer with L1-bypassing 128-bit loads)
        import ctypes
        lib = L.load()
        buf = torch.empty(48 * 1024 * 1024 // 4, dtype=torch.float32, device=dev).normal_()
        sink = torch.zeros(1, device=dev)
        probe = lambda it: lib.cwt_debug_l2_read(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 4, it, n_sm * 4,
                                                 ctypes.c_void_p(sink.data_ptr()), L.stream_ptr(dev))
        l2_ms, _ = timed(lambda: probe(20))
        l2_peak = buf.numel() * 4 * 20 / (l2_ms / 1e3) / 1e9
        del buf
        roofline = {
            "kernel": "k_fit_l2 (persistent cooperative kernel: two 5-shot episodes at a time re-laid out tile by tile and kept "
                      "L2-resident; every sweep streams them through a 3-stage shared-memory ring with 51.2 KB bulk-TMA copies)",
            "bound": "l2", "achieved": fit_gbs, "peak": l2_peak, "unit": "GB/s", "frac": fit_gbs / l2_peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "L2 read bandwidth measured in this run (cwt_debug_l2_read, 48 MB working set, 20 sweeps)",
            "note": "achieved = algorithmic bytes (2T+1)*S*F per episode / CUDA-event time of the fit call (SURVEY §8d); "
                    "the same bytes / the measured HBM peak is hbm_equivalent",
            "hbm_equivalent": {"achieved": fit_gbs, "peak": hbm_peak, "x_of_hbm_peak": fit_gbs / hbm_peak, "peak_source": peak_src},
            "dram_frac": (traffic / (fit_ms / 1e3) / 1e9 / hbm_peak) if traffic else None,
            "compulsory_dram_bytes": E * S * F_bytes,
        }
    else:
        roofline = {
            "kernel": "fit_classifier streaming: 200 x {rows_times_feat<1>, fit_hires, feat_times_cols<1>+SGD}",
            "bound": "hbm", "achieved": fit_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fit_gbs / hbm_peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "note": "achieved = algorithmic bytes (2T+1)*S*F per episode / CUDA-event time of the fit call (SURVEY §8d)",
        }
    roofline.update({
        "algorithmic_bytes_per_call": bytes_fit, "ms_per_call": fit_ms,
        "stages_ms": {"fit": fit_ms, "transformer": tr_ms, "logits_iou": iou_ms},
        "logits_iou": {"bound": "hbm", "achieved": bytes_iou / (iou_ms / 1e3) / 1e9, "frac": bytes_iou / (iou_ms / 1e3) / 1e9 / hbm_peak},
        "transformer": {"bound": "hbm", "achieved": bytes_tr / (tr_ms / 1e3) / 1e9, "frac": bytes_tr / (tr_ms / 1e3) / 1e9 / hbm_peak,
                        "note": "algorithmic bytes count f_q once; the re-associated attention needs two passes over it"},
    })
    if a.attn_both:
        tc_ms, w_tc = timed(lambda: cwt.transformer_forward(w_fit, devb.f_q, params["w_qkvs.weight"], params["fc.weight"],
                                                            params["fc.bias"], params["layer_norm.weight"],
                                                            params["layer_norm.bias"], a.heads, normalize_k=True, algo=1))
        flops = E * a.heads * 2.0 * 3600 * 512 * 512 * 3                  # 3 bf16 products (hi.hi + hi.lo + lo.hi) per fp32 product
        tpeak = float(peaks.get("bf16_tflops", 1590.0))
        roofline["transformer_tcgen05"] = {
            "bound": "tensor", "ms_per_call": tc_ms, "achieved": flops / (tc_ms / 1e3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
            "frac": flops / (tc_ms / 1e3) / 1e12 / tpeak,
            "max_rel_diff_vs_reassoc": float((w_tc - w_ad).norm() / w_ad.norm()),
            "note": "whole block with the K projection as a tcgen05 GEMM (3 x bf16 split); flops = E*nH*2*HW*C*C*3 over the time of "
                    "the WHOLE block (pre-pass, GEMM, softmax, V side, fc, LayerNorm); the re-associated path (default) needs 127x fewer flops"}

    # ---- end to end through the public API with host buffers ("e2e") ----
    # few_shot_seg_cwt_b200.HostPipeline: every step copies its inputs from pinned host memory (side stream, overlapping
    # the previous step's head) and reads the step's int64 counts back to the host.
    e2e = None
    if not a.no_e2e:
        def run_e2e(host_batch, tag):
            pipe = cwt.HostPipeline(dev, params, a.heads, a.cls_lr, a.adapt_iter, fit_algo=a.fit_algo, attn_algo=a.attn_algo,
                                    sub_batch=a.e2e_sub_batch, sub_batch_all=bool(a.e2e_sub_all), expand_on_main=int(a.e2e_expand_main),
                                    ramp=tuple(int(x) for x in a.e2e_ramp.split(",") if x))
            pipe.run([host_batch] * max(1, a.warmup), reduce_every_step=True)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = pipe.run([host_batch] * a.steps, reduce_every_step=True)
            e1.record()
            barrier()
            ms_e = max_over_ranks(e0.elapsed_time(e1))
            return {"value": world * E * a.steps / (ms_e / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": host_batch.nbytes() + host.subcls.numel() * 16,      # + subcls and idx (int64)
                    "d2h_bytes_per_step": res[0].numel() * res[0].element_size() + 4 * E, "ms_per_step": ms_e / a.steps,
                    "host_format": tag,
                    "counts_equal_device_resident_run": bool(torch.equal(res[0].to(dev), last_out.counts))}
        api = ("few_shot_seg_cwt_b200.HostPipeline.run (four device staging slots; first batch of the run in sub-batches of "
               "%s episodes, %s; expansion of the zero-compressed features on %s; async D2H of the counts and the fit status words)"
               % (a.e2e_ramp or a.e2e_sub_batch, "every batch in sub-batches of %d" % a.e2e_sub_batch if a.e2e_sub_all else "later batches whole",
                  {0: "the copy stream", 1: "the head's stream", 2: "its own stream (under the previous piece's fit)"}.get(int(a.e2e_expand_main))))
        dense = run_e2e(host, "dense fp32 tensors in pinned host memory")
        if a.e2e_format == "dense":
            e2e = dict(dense, api=api)
        else:
            # lossless zero-compressed transport of the post-ReLU features (few_shot_seg_cwt_b200.hostformat): bit mask + prefix
            # counts + packed non-zeros cross PCIe, cwt_expand_zero_compressed_f32 rebuilds the dense tensors on the device
            # inside the timed region; the dense-format number is kept beside it
            from few_shot_seg_cwt_b200 import hostformat
            comp = hostformat.compress_batch(host).pin_memory()
            e2e = dict(run_e2e(comp, "zero-compressed features (bit mask + prefix counts + packed non-zeros), expanded on the "
                                     "device inside the timed region; labels / weights dense"), api=api)
            e2e["dense_format"] = {k: dense[k] for k in ("value", "h2d_bytes_per_step", "ms_per_step", "counts_equal_device_resident_run")}

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        # the CPU leg runs the oracle on the first episodes of THIS rank's batch (outside every timed region); its result for
        # episode 0 doubles as the parity check of the numbers above
        rate, dt, ora, kind, desc = cpu_reference_rate(a, a.cpu_baseline_episodes, warm=1, first_index=int(host.idx[0]))
        cpu = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{a.cpu_baseline_episodes} episodes of the same workload after 1 warm-up ({dt:.1f} s), "
                         f"{desc}, torch {torch.__version__} CPU fp32"}
        got = last_out.counts[0].cpu()
        rel = lambda x, y: float((x.double().cpu() - y.double()).norm() / y.double().norm())
        scale = max(1.0, float(ora["logits60"].abs().max()))
        tie = int((ora["tie_margin"] <= 1e-5 * scale).sum())
        tie0 = int((ora["tie_margin0"] <= 1e-5 * max(1.0, float(ora["logits60_0"].abs().max()))).sum())
        d_ad = int((got[0] - ora["counts"]).abs().max())
        d_bl = int((got[1] - ora["counts0"]).abs().max())
        parity = {"episode": int(host.idx[0]), "w_fit_rel_err": rel(last_out.w_fit[0], ora["W_fit"]),
                  "w_adapted_rel_err": rel(last_out.w_adapted[0], ora["W_adapted"]),
                  "counts_max_abs_diff": {"adapted": d_ad, "baseline": d_bl}, "tie_set_pixels": {"adapted": tie, "baseline": tie0},
                  "ok": bool(rel(last_out.w_fit[0], ora["W_fit"]) < 1e-4 and rel(last_out.w_adapted[0], ora["W_adapted"]) < 1e-4
                             and d_ad <= tie and d_bl <= tie0),
                  "against": f"{desc} on the same synthetic episode (checked outside the timed regions)"}

    if rank == 0:
        emit(json.dumps({
            "metric": metric_name(a), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "episodes_per_gpu_per_step": E, "global_episodes_per_step": E * world,
                       "parallelism": f"episodes sharded over {world} GPU(s), int64 IoU all-reduce",
                       "pipelining": "post stage of step i (transformer, logits/IoU, all-reduce) on a side stream under the fit of step i+1",
                       "l2_policy": f"inputs larger than L2 ({host.nbytes() / 1e6:.0f} MB per step per GPU, streamed every SGD step)",
                       "distinct_episodes_per_gpu": nd,
                       "fit_algo": a.fit_algo,
                       "attn_algo": f"{a.attn_algo} ({'re-associated scores on CUDA cores (default: 127x fewer flops)' if a.attn_algo == 0 else 'K projection as a tcgen05/TMEM GEMM'})"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "parity_check": parity, "miou_adapted": table.miou(0), "miou_baseline": table.miou(1),
            "bad_episodes": int(table.n_bad),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
