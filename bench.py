#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric: codec tokens/s at 830M decode).

Workload (BASELINE.json configs[1], SURVEY.md section 8d "Config 2"): giga830M shape (d=2048, 16 heads, 16 layers,
K=4 codebooks, vocab 2048+4), random-init bf16-representable weights, B=32 independent synthetic utterances per
GPU (80 phoneme ids, 150-frame / 3 s prompt -> up to 800 frames / 16 s), end tokens suppressed so only the
reference's length cap stops generation.  A "step" is ONE decode step of the whole batch: B frames = B*K codec
tokens, every layer + logit heads + the fused sampler.

    python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference ...                            the CPU reference arm (oracle port, all host threads)

value    = N * B * K * steps / device time (CUDA events, max over ranks); inputs resident in HBM.
e2e      = same metric through the public API (VoiceCraft.inference_tts_many) with pinned HOST inputs and host
           outputs: prefill + every decode step + H2D/D2H inside the timed region.
roofline = dominant kernel (by device time in a profiled pass of the same steps): algorithmic bytes / launch over its
           CUDA-event duration, against MEASURED_PEAKS.json; step_roofline = same for the whole step.
"""
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="830M")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--codebooks", type=int, default=4)
    ap.add_argument("--text-len", type=int, default=80)
    ap.add_argument("--prompt", type=int, default=150)
    ap.add_argument("--kv", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-steps", type=int, default=4, help="decode steps of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_model_inputs(args, device=None):
    from voicecraft_b200 import synthetic
    over = {}
    if args.codebooks != 4:
        over["n_codebooks"] = args.codebooks
    cfg = synthetic.make_config(args.model, **over)
    sd = synthetic.make_state_dict(cfg, seed=0)
    end = cfg.eos if cfg.eos > 0 else cfg.eog
    for k in range(cfg.n_codebooks):            # only the length cap ends generation
        sd[f"predict_layer.{k}.2.bias"][end] = -1e4
        sd[f"predict_layer.{k}.2.bias"][cfg.eog] = -1e4
    utts = [synthetic.synthetic_utterance(cfg, 100 + i, args.text_len, args.prompt) for i in range(args.batch)]
    return cfg, sd, utts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_bytes(cfg, B, S_mean, kv_bytes):
    """bf16 weights + KV read (B*S tokens) + KV write (B tokens), per decode step."""
    d, L, K = cfg.d_model, cfg.num_decoder_layers, cfg.n_codebooks
    V = 2048 + cfg.n_special
    per_layer = 3 * d * d + d * d + 8 * d * d
    heads = K * ((1024 * d) + V * 1024)
    W = 2 * (L * per_layer + heads)
    kv_tok = L * 2 * d * kv_bytes
    return W, B * S_mean * kv_tok, B * kv_tok


def host_threads():
    """Usable host cores: affinity mask, cgroup CPU quota, capped at 64 (fp32 GEMV-like steps stop scaling earlier)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        pass
    env = os.environ.get("BENCH_CPU_THREADS")
    return int(env) if env else max(1, min(n, 64))


def cpu_baseline(args, cfg, sd, utts, steps, threads=None):
    """The oracle port of the reference's own batched decode (inference_tts_batch, B copies of one prompt ==
    the compute of B independent utterances of that length) timed on the host cores, bounded sample."""
    from oracle import lm_oracle
    threads = threads or host_threads()
    torch.set_num_threads(threads)
    oracle = lm_oracle.OracleLM(cfg, sd)
    x, x_lens, y = utts[0]
    marks = []
    torch.manual_seed(1)
    oracle.inference_tts_batch(x, x_lens, y, top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3,
                               batch_size=args.batch, max_steps=steps + 1, on_step=lambda c: marks.append(time.perf_counter()))
    dt = marks[-1] - marks[0]                      # decode steps only (the first mark is after prefill + first sample)
    n = len(marks) - 1
    tok_s = args.batch * cfg.n_codebooks * n / dt
    return dict(value=tok_s, unit="codec tokens/s", cores=threads, kind="port",
                sample=f"oracle inference_tts_batch B={args.batch}, {n} decode steps after a {args.text_len + args.prompt + 1}-token "
                       f"prefill, fp32, {threads} threads, {dt / n * 1e3:.0f} ms/step"), dt / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, sd, utts = make_model_inputs(args)
    steps = max(1, min(args.steps, 64))
    warm = max(0, min(args.warmup, 2))
    cb, ms = cpu_baseline(args, cfg, sd, utts, steps + warm)
    line = {"impl": "reference", "metric": "codec tokens/s (830M TTS decode)", "value": cb["value"], "unit": "codec tokens/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "codec tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, cfg):
    return {"workload": f"giga{args.model} TTS decode, B={args.batch}/GPU independent utterances, K={cfg.n_codebooks}, "
                        f"text {args.text_len}, prompt {args.prompt} frames, ctx {args.text_len + args.prompt + 1}+",
            "batch_per_gpu": args.batch, "n_codebooks": cfg.n_codebooks, "kv_cache": args.kv,
            "l2": "per-step working set (1.65 GB weights + >=0.9 GB KV) >> 126 MB L2: no flush needed",
            "sampling": "top_k=40, top_p=1.0, temperature=1.0, one [B*K,V] Exp(1) draw per step"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from voicecraft_b200 import _lib
    from voicecraft_b200.voicecraft import VoiceCraft
    import ctypes as C

    cfg, sd, utts = make_model_inputs(args)
    K, B = cfg.n_codebooks, args.batch
    model = VoiceCraft(cfg)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    cap = args.text_len * (cfg.encodec_sr // 5)
    max_steps_avail = cap - (args.prompt + 1)            # decode steps until the length cap fires
    W = max(3, args.warmup)
    Ksteps = min(args.steps, max_steps_avail - W - 2)
    model.configure_engine(max_slots=B, max_seq_len=(args.text_len + cap + 64 + 255) // 256 * 256, kv_dtype=args.kv,
                           max_new_tokens=cap + 64)
    xs = [u[0] for u in utts]
    ys = [u[2] for u in utts]
    kw = dict(top_k=40, top_p=1.0, temperature=1.0, stop_repetition=3)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: device-timed decode steps
    torch.manual_seed(1 + rank)
    sess = model.open_tts_session([x.to(dev) for x in xs], [y.to(dev) for y in ys], **kw)
    eng = sess.eng
    sess.sample()
    for _ in range(W):
        sess.step()
    ctx0 = args.text_len + args.prompt + 1 + W
    clocks = ClockSampler(local)
    launches0 = lib.vcb_counter(eng, b"launches")
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(Ksteps):
        sess.step()
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    launches = lib.vcb_counter(eng, b"launches") - launches0
    chain_info = {"clusters": int(lib.vcb_counter(eng, b"chain_clusters")), "epoch": int(lib.vcb_counter(eng, b"chain_epoch"))}
    st = sess.poll()
    assert all(s.n_steps == 1 + W + Ksteps for s in st), [s.n_steps for s in st]
    assert not any(s.done for s in st)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tok_s = world * B * K * Ksteps / (ms * 1e-3)
    ctx1 = ctx0 + Ksteps

    # ------------------------------------------------------------------ roofline: profiled pass (same engine state)
    roof = None
    step_roof = None
    if rank == 0:
        peak, peak_src = peaks()
        kvb = 4 if args.kv == "fp32" else 2
        S_mean = (ctx0 + ctx1) / 2.0
        Wb, KVr, KVw = algorithmic_bytes(cfg, B, S_mean, kvb)
        step_bytes = Wb + KVr + KVw
        step_gbs = step_bytes / (ms / Ksteps * 1e-3) / 1e9
        step_roof = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                     "algorithmic_bytes_per_step": step_bytes, "weights_bytes": Wb, "kv_read_bytes": KVr,
                     "peak_source": peak_src, "ctx_mean": S_mean}
        # profile a few more steps with events around every launch
        nprof = min(8, max_steps_avail - (W + Ksteps) - 2)
        if nprof > 0:
            lib.vcb_set_option(eng, b"profile", 1)
            for _ in range(nprof):
                sess.step()
            msb = (C.c_double * 7)()
            cnt = (C.c_int64 * 7)()
            lib.vcb_profile_read(eng, msb, cnt, 7)
            lib.vcb_set_option(eng, b"profile", 0)
            names = ["gemm_w_xT_cluster(tcgen05, cluster split-K)", "attn_rows_kernel(paged KV, TMA bulk, split ctx)",
                     "ln_rows_kernel", "(unused)", "sampler_kernel", "step_prep", "mega_step_kernel(persistent decode step)"]
            total = sum(msb)
            shares = {names[i]: {"ms_per_step": msb[i] / nprof, "launches_per_step": cnt[i] / nprof, "share": msb[i] / total}
                      for i in range(7)}
            S_prof = ctx1 + nprof / 2.0
            L = cfg.num_decoder_layers
            dom = max(range(7), key=lambda i: msb[i])
            if dom == 6:
                bytes_per_launch = sum(algorithmic_bytes(cfg, B, S_prof, kvb))     # the whole step is one launch
            elif dom == 1:
                bytes_per_launch = B * (S_prof + 1) * 2 * cfg.d_model * kvb      # K+V rows of every cached token, one layer
            else:
                bytes_per_launch = Wb / (cnt[0] / nprof)                          # mean weight bytes per GEMM launch
            # Average launch duration over the TIMED region: the kernel's share of the step (from the event-per-launch
            # pass, which serialises launches and adds event overhead to each, so only the share is used) applied to
            # the measured step time.  The isolated per-launch event time is reported next to it.
            iso = msb[dom] / cnt[dom] * 1e-3
            dur = (msb[dom] / total) * (ms / Ksteps * 1e-3) / (cnt[dom] / nprof)
            ach = bytes_per_launch / dur / 1e9
            # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/r01_ncu_summary.md): the GEMMs
            # move exactly their weight bytes, attention 1.09x its algorithmic KV bytes
            traffic = bytes_per_launch * (1.09 if dom == 1 else 1.0)
            roof = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": "profiles/r01_ncu_summary_v5.md (dram__bytes_read+write per launch / algorithmic)", "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_us": dur * 1e6, "isolated_launch_us": iso * 1e6,
                    "peak_source": peak_src, "ctx": S_prof, "by_kernel": shares,
                    "note": "avg_launch_us = share of the step (event-per-launch pass) x timed step / launches; isolated_launch_us = raw per-launch event time (serialised, no PDL overlap)"}
    sess.close()

    # ------------------------------------------------------------------ e2e: public API, host in / host out
    e2e = None
    if not args.no_e2e:
        xs_h = [x.pin_memory() for x in xs]
        ys_h = [y.pin_memory() for y in ys]
        torch.manual_seed(1 + rank)
        barrier()
        t0 = time.perf_counter()
        out = model.inference_tts_many(xs_h, ys_h, poll_every=8, **kw)
        res_h = [r[0].cpu() for r in out]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        gen_frames = sum(int(r[1].shape[-1]) for r in out)
        steps_e2e = model.last_stats.get("steps", 0) or (gen_frames // B + K)
        h2d = sum(x.numel() * 8 for x in xs_h) + sum(y.numel() * 8 for y in ys_h)
        d2h = sum(r.numel() * 8 for r in res_h)
        e2e = {"value": world * gen_frames * K / dt, "unit": "codec tokens/s", "h2d_bytes_per_step": h2d / max(steps_e2e, 1),
               "d2h_bytes_per_step": d2h / max(steps_e2e, 1), "seconds": dt, "generated_frames": gen_frames,
               "note": "prefill + all decode steps + polling + H2D of prompts + D2H of tokens inside the timed region"}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cb = None
    if not args.no_cpu:
        cb, _ = cpu_baseline(args, cfg, sd, utts, args.cpu_steps)
    frames_s = tok_s / K
    line = {"metric": "codec tokens/s (830M TTS decode)", "value": tok_s, "unit": "codec tokens/s", "n_gpus": world,
            "steps": Ksteps, "warmup": W, "ms_per_step": ms / Ksteps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": dict(workload_config(args, cfg), ctx_start=ctx0, ctx_end=ctx1),
            "rtf_per_stream": frames_s / (world * B) / cfg.encodec_sr, "frames_per_s": frames_s,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "gemm_chain": chain_info, "roofline": roof, "step_roofline": step_roof,
            "cpu_baseline": cb}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
