#!/usr/bin/env python
"""bench.py -- headline benchmark of the lm.rs hot path on B200 (BASELINE.json: decode tok/s + prefill tok/s,
Llama-3.2-1B Q8_0, vs the CPU path, with the HBM roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model NAME] [--pos P]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

A "step" is one decode pass (Transformer::forward, src/transformer.rs:316) of one token at position >= P over a
synthetic LMRS file with the real model's shapes (no weights/tokenizers exist offline).  One JSON line:
  value     HBM-resident decode throughput: K forward steps enqueued back to back on the device (token ids and the
            KV cache already in HBM), timed with CUDA events on the launching stream, max over ranks.
  e2e       the same metric through the public API in a greedy generate loop like src/bin/chat.rs:188-226 with
            --temperature 0.  Headline: `forward_argmax(token, pos) -> next token` (forward + Sampler::sample_argmax fused
            on the device, src/sampler.rs:29-41): per step 16 B of step parameters go host->device and the 4-byte token id
            comes back.  `e2e.logits_to_host` is the same loop through `forward(token, pos) -> host logits` (vocab*4 bytes
            device->host per step + numpy argmax), `e2e.generate_greedy` the whole loop in ONE call with the token fed
            back on the device.
  parity    the timed workload replayed on the CPU oracle in the same run: the residual stream returned by the
            512-embedding fill_kv_cache and the logits of the first greedy decode steps at pos 512+, max-abs difference and
            bit-exactness.  (The oracle is a restatement, the Rust binary cannot be built here: parity is unpinned.)
  prefill   fill_kv_cache(P embeddings) (src/transformer.rs:672) timed end to end through the C ABI.
  roofline  decode is HBM-bound: algorithmic bytes per step (lmrs_file.decode_bytes_per_token: every weight byte and
            scale once + KV rows) / step time vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle (C restatement of the reference, "port") timed on this box's host cores on a bounded
            sample of the same workload.
--impl reference runs ONLY that CPU path (rank 0) and prints the same line with "impl": "reference".
Inputs are larger than L2 (1.27 GB of weights streamed per step vs 126 MB L2), so no explicit L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the process group of this benchmark only carries barriers and the max-reduce of the timings: no NVLink-SHARP multicast
# buffers to set up and tear down (the decode data path does not use NCCL at all)
os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402


def load_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_run(buf, a, pos0, steps, warmup, prompt, keep_logits=0, kshards=1):
    """The reference arm / cpu_baseline: oracle/ (C restatement of the reference's CPU path) on the host cores.
    keep_logits: also return the prefill residual stream and the logits of the first `keep_logits` greedy steps from
    (prompt[pos0], pos0) -- the parity reference for the GPU arm."""
    import lmrs_ref
    lmrs_ref.build()
    lmrs_ref.set_kshards(kshards)   # N-GPU parity leg: the rank-ordered partial sums of the peer exchange (oracle/lmrs_ref.h)
    m = lmrs_ref.RefTransformer(buf)
    # "all the host threads it can use": probe a few team sizes on one decode step each and keep the fastest
    # (nproc can exceed the container's CPU quota, where more threads only add spin-wait contention)
    cores = lmrs_ref.usable_cores()
    best, best_t = None, None
    for n in sorted({c for c in (cores, cores // 2, 32, 16, 8) if 1 <= c <= cores}, reverse=True):
        lmrs_ref.set_threads(n)
        m.forward(int(prompt[0]), 0)
        t0 = time.perf_counter()
        m.forward(int(prompt[1]), 1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    lmrs_ref.set_threads(best)
    emb = m.get_embeddings(prompt[:pos0])
    t0 = time.perf_counter()
    m.fill_kv_cache(emb, 0)
    t_prefill = time.perf_counter() - t0
    kept = []
    tok = int(prompt[pos0])
    for i in range(keep_logits):      # (rows >= pos are never read: the warm-up / timed steps below simply overwrite them)
        lg = m.forward(tok, pos0 + i).copy()
        kept.append(lg)
        tok = int(np.argmax(lg))
    tok = int(prompt[pos0])
    for i in range(warmup):
        tok = int(np.argmax(m.forward(tok, pos0 + i)))
    t0 = time.perf_counter()
    for i in range(steps):
        tok = int(np.argmax(m.forward(tok, pos0 + warmup + i)))
    dt = time.perf_counter() - t0
    return {"decode_tok_s": steps / dt, "ms_per_step": dt / steps * 1e3, "prefill_tok_s": pos0 / t_prefill,
            "cores": lmrs_ref.lib().lmrs_ref_num_threads(), "steps": steps, "prefill_stream": emb, "logits": kept}


def load_gemv_traffic(model, quant):
    """dram__bytes_read.sum + dram__bytes_write.sum per gemv_kernel launch (average over a step's launches) from the
    committed `ncu --set full` capture of this workload (profiles/r2_gemv_traffic.json); None for workloads not captured."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r2_gemv_traffic.json")
    try:
        d = json.load(open(p))
    except OSError:
        return None, "no ncu capture committed"
    w = d.get("workload", {})
    if w.get("model") != model or w.get("quant") != quant:
        return None, "no ncu --set full capture for this workload"
    return d["traffic_bytes_per_launch_avg"], "profiles/r2_gemv_traffic.json (ncu --set full, per-launch average)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="llama-3.2-1b")
    ap.add_argument("--quant", type=int, default=1, help="1 = Q8_0, 2 = Q4_0")
    ap.add_argument("--pos", type=int, default=512, help="prompt length / first decode position")
    ap.add_argument("--cpu-steps", type=int, default=16, help="decode steps of the bounded cpu_baseline sample")
    ap.add_argument("--parity-steps", type=int, default=4, help="greedy decode steps replayed on the CPU oracle (0: skip the parity block)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from lmrs_b200 import lmrs_file as lf
    a = lf.model_args(args.model, args.quant)
    metric = f"decode tok/s {args.model} {'Q8_0' if args.quant == 1 else 'Q4_0'} (prefill tok/s in 'prefill')"
    config = {"workload": f"{args.model} {'Q8_0' if args.quant == 1 else 'Q4_0'} synthetic LMRS v4 file: greedy decode of "
                          f"{args.steps} tokens from pos {args.pos} after a {args.pos}-embedding fill_kv_cache",
              "model_file_bytes": lf.file_size(a), "first_pos": args.pos, "batch": 1,
              "l2": "inputs larger than L2 (whole weight set streamed every step); no flush needed",
              "parallelism": "single GPU" if args.gpus == 1 else
                             f"row-sharded x{args.gpus}: partial Wo/W2 results pushed between the GPUs by the kernels (NVLink peer stores, "
                             f"rank-ordered sum in the next prologue), no collective in the decode chain"
                             if os.environ.get("LMRS_B200_PEER", "1") != "0" else f"row-sharded x{args.gpus} (2 NCCL all-reduces per block)"}

    if args.impl == "reference":
        if rank != 0:
            return
        buf = lf.write_synthetic(a, seed=0, mode="fast")
        prompt = np.random.default_rng(1).integers(0, a.vocab_size, args.pos + 1).astype(np.uint32)
        r = cpu_reference_run(buf, a, args.pos, args.steps, args.warmup, prompt)
        sample = f"{args.steps} greedy decode steps at pos {args.pos}+ after a {args.pos}-token batched prefill"
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": r["decode_tok_s"], "unit": "tok/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int8 x int8 -> int32 groups, f32 accumulate", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port", "sample": sample,
                             "prefill_tok_s": r["prefill_tok_s"],
                             "note": "C/OpenMP restatement of the reference (oracle/); the Rust/rayon binary cannot be built here (no rustc)"},
            "e2e": {"value": r["decode_tok_s"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "prefill": {"value": r["prefill_tok_s"], "unit": "tok/s", "tokens": args.pos}}))
        return

    import torch
    import lmrs_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (lmrs_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    buf = lf.write_synthetic(a, seed=0, mode="fast")
    prompt = np.random.default_rng(1).integers(0, a.vocab_size, args.pos + 1).astype(np.uint32)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(lmrs_b200.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        m, _ = lmrs_b200.Transformer.new_sharded(buf, local_rank, rank, world, bytes(idt.cpu().numpy().tobytes()))
    else:
        m, _ = lmrs_b200.Transformer.new(buf, local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank); sampler.start()   # nvidia-smi needs ~100 ms to come up: started before the prefill,
                                                           # samples taken before the timed decode region are dropped below
    # ---- prefill: fill_kv_cache(P embeddings), end to end through the C ABI ------------------------------------
    emb0 = m.get_embeddings(prompt[:args.pos])
    emb = emb0.copy()
    m.fill_kv_cache(emb, 0)            # warm-up: first call allocates the batch buffers and loads the kernels
    emb = emb0.copy()                  # (idempotent: the same rows are written to the same cache positions)
    barrier()
    t0 = time.perf_counter()
    newpos = m.fill_kv_cache(emb, 0)
    t_prefill = time.perf_counter() - t0
    prefill_dev_ms = m.last_prefill_device_ms()
    assert newpos == args.pos
    # ---- parity leg (GPU side): greedy steps from (prompt[P], P), logits kept; compared with the oracle further down ----
    gpu_logits = []
    tok = int(prompt[args.pos])
    for i in range(args.parity_steps):
        lg = m.forward(tok, args.pos + i).copy()
        gpu_logits.append(lg)
        tok = int(np.argmax(lg))

    # ---- decode, HBM-resident leg ("value") --------------------------------------------------------------------
    stream = torch.cuda.Stream()       # a real (non-default) stream: the library launches on it, the events time it
    assert stream.cuda_stream != 0
    m.set_stream(stream.cuda_stream)
    toks = np.random.default_rng(2).integers(0, a.vocab_size, args.warmup + args.steps)
    pos = args.pos
    for i in range(args.warmup):       # warm-up and timed steps cover the SAME positions P..P+K-1 (BASELINE: "64 tokens at pos 512")
        m.forward_device(int(toks[i]), pos + (i % args.steps))
    barrier()
    l0 = m.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.rows.clear()               # keep only samples from the timed region (+ the e2e leg)
    ev0.record(stream)
    for i in range(args.steps):
        m.forward_device(int(toks[args.warmup + i]), pos); pos += 1
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = m.kernel_launches() - l0
    # ---- the dominant, HBM-bound kernel by itself: only the lmrs_q_matvec_kernel launches of a step (no attention) -----------
    for i in range(args.warmup):
        n_gemv = m.bench_gemv_pass(args.pos + i)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for i in range(args.steps):
        m.bench_gemv_pass(args.pos + i)
    g1.record(stream)
    barrier()
    gemv_ms = g0.elapsed_time(g1) / args.steps
    m.set_stream(0)
    if dist is not None:
        t = torch.tensor([dev_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = 1e3 / ms_per_step

    # ---- decode, end-to-end legs (chat.rs generate loop, temperature 0) ---------------------------------------------
    def timed_loop(step_fn):
        tok = int(prompt[args.pos])
        for i in range(args.warmup):
            tok = step_fn(tok, args.pos + (i % args.steps))
        barrier()
        t0 = time.perf_counter()
        pos = args.pos
        for i in range(args.steps):
            tok = step_fn(tok, pos); pos += 1
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        return dt
    # (a) forward() -> host logits -> numpy argmax: vocab*4 bytes device->host per step
    e2e_logits_s = timed_loop(lambda tok, pos: int(np.argmax(m.forward(tok, pos))) % a.vocab_size)
    peer = os.environ.get("LMRS_B200_PEER", "1") != "0"
    if world == 1 or peer:
        # (b) headline: forward_argmax() -- the greedy pick happens on the device, 4 bytes come back
        e2e_s = timed_loop(lambda tok, pos: m.forward_argmax(tok, pos))
        # (c) the whole loop in one call, token fed back on the device
        m.generate_greedy(int(prompt[args.pos]), args.pos, min(args.warmup, args.steps))
        t0 = time.perf_counter()
        gen = m.generate_greedy(int(prompt[args.pos]), args.pos, args.steps)
        gen_s = time.perf_counter() - t0
        assert len(gen) == args.steps
    else:
        e2e_s, gen_s = e2e_logits_s, None      # the fused sampler is single-GPU (sharded logits are gathered to every rank)
    e2e = args.steps / e2e_s
    clocks = sampler.stop()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            m.close()
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    peak *= args.gpus
    mid_pos = args.pos + args.steps // 2
    alg_bytes = lf.decode_bytes_per_token(a, mid_pos)
    w_bytes = lf.decode_bytes_per_token(a)                      # weights + scales + norm vectors: what lmrs_q_matvec_kernel streams
    achieved = (w_bytes / n_gemv) / (gemv_ms / n_gemv * 1e-3) / 1e9
    step_gbs = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic, traffic_src = load_gemv_traffic(args.model, args.quant)
    roofline = {"bound": "hbm", "kernel": f"lmrs_q_matvec_kernel ({n_gemv} launches per step: 4 per block + classifier)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src + (f" x {args.gpus} GPUs" if args.gpus > 1 else ""),
                "algorithmic_bytes_per_launch": w_bytes / n_gemv, "avg_launch_us": gemv_ms / n_gemv * 1e3,
                "how": "CUDA events around the step's lmrs_q_matvec_kernel launches alone (PDL-chained, real prologues/epilogues, attention skipped)",
                "traffic": traffic, "traffic_source": traffic_src,
                "whole_step": {"algorithmic_bytes_per_step": alg_bytes, "achieved": step_gbs, "frac": step_gbs / peak,
                               "launches_per_step": launches / args.steps,
                               "note": "includes the latency-bound exact-order attention (not an HBM-bound kernel)"}}
    cpu = None
    parity = None
    if (args.gpus == 1 and args.cpu_steps > 0) or args.parity_steps > 0:
        cpu_steps = args.cpu_steps if args.gpus == 1 else 0
        r = cpu_reference_run(buf, a, args.pos, max(cpu_steps, 1), 2 if cpu_steps else 0, prompt, keep_logits=args.parity_steps,
                              kshards=args.gpus if (args.gpus > 1 and peer) else 1)
        if cpu_steps:
            cpu = {"value": r["decode_tok_s"], "unit": "tok/s", "cores": r["cores"], "kind": "port",
                   "sample": f"{args.cpu_steps} greedy decode steps at pos {args.pos}+ after a {args.pos}-token batched prefill",
                   "prefill_tok_s": r["prefill_tok_s"]}
        if args.parity_steps > 0:
            d_stream = float(np.abs(emb - r["prefill_stream"]).max())
            d_logits = [float(np.abs(g - c).max()) for g, c in zip(gpu_logits, r["logits"])]
            parity = {"reference": "oracle/ (C restatement of the reference's CPU path; the Rust binary cannot be built here: parity unpinned)"
                                   + (f"; N-GPU: Wo/W2 partial sums added in rank order (oracle k-shard mode, {args.gpus} shards)" if args.gpus > 1 and peer else ""),
                      "prefill_residual_stream": {"rows": args.pos, "max_abs": d_stream, "bit_exact": bool(np.array_equal(emb, r["prefill_stream"]))},
                      "decode_logits": {"positions": [args.pos + i for i in range(len(d_logits))], "max_abs": max(d_logits) if d_logits else None,
                                        "bit_exact": bool(all(np.array_equal(g, c) for g, c in zip(gpu_logits, r["logits"]))),
                                        "greedy_tokens_equal": bool(all(int(np.argmax(g)) == int(np.argmax(c)) for g, c in zip(gpu_logits, r["logits"])))},
                      "max_abs": max([d_stream] + d_logits), "tolerance": 1e-3,
                      "bit_exact": bool(np.array_equal(emb, r["prefill_stream"]) and all(np.array_equal(g, c) for g, c in zip(gpu_logits, r["logits"])))}
    print(json.dumps({
        "metric": metric, "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int8 x int8 -> int32 groups, f32 accumulate", "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 4 if (world == 1 or peer) else a.vocab_size * 4,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "api": "forward_argmax(token, pos) -> next token (greedy pick fused on the device)" if (world == 1 or peer) else "forward(token, pos) -> host logits",
                "logits_to_host": {"value": args.steps / e2e_logits_s, "ms_per_step": e2e_logits_s / args.steps * 1e3,
                                   "d2h_bytes_per_step": a.vocab_size * 4, "api": "forward(token, pos) -> host logits + host argmax"},
                "generate_greedy": None if gen_s is None else {"value": args.steps / gen_s, "ms_per_step": gen_s / args.steps * 1e3,
                                                                "api": "generate_greedy(first_token, pos, n): one call, token fed back on the device"}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        "prefill": {"value": args.pos / t_prefill, "unit": "tok/s", "tokens": args.pos, "ms": t_prefill * 1e3,
                    "device_ms": prefill_dev_ms, "device_tok_s": args.pos / (prefill_dev_ms * 1e-3) if prefill_dev_ms > 0 else None,
                    "h2d_bytes": int(emb.nbytes), "d2h_bytes": int(emb.nbytes),
                    "path": "fill_kv_cache end to end through the C ABI (host embeddings in, residual stream out)",
                    "roofline": {"bound": "tensor", "unit": "TOP/s", "achieved": lf.prefill_int8_ops(a, args.pos) / (prefill_dev_ms * 1e-3) / 1e12,
                                 "peak": 4500.0, "peak_source": "nominal dense int8 (B200_PROFILING.md); MEASURED_PEAKS.json has no int8 figure",
                                 "frac": lf.prefill_int8_ops(a, args.pos) / (prefill_dev_ms * 1e-3) / 1e12 / 4500.0,
                                 "end_to_end_frac": lf.prefill_int8_ops(a, args.pos) / t_prefill / 1e12 / 4500.0,
                                 "note": "device time of the whole fill_kv_cache (GEMMs + exact-order attention + row kernels, CUDA events) over the matmul ops only; end_to_end_frac includes the PCIe copies"}},
        "published_reference": {"value": 50, "unit": "tok/s", "hardware": "16-core AMD Epyc (README.md:38)"} if args.model == "llama-3.2-1b" and args.quant == 1 else None,
    }))
    if dist is not None:
        dist.barrier()
        m.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
