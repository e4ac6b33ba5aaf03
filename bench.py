#!/usr/bin/env python
"""Benchmark of the MusicTransformer hot path (BASELINE.json metric: train tokens/s @ L=2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config B|A|C]

One "step" = one optimizer step of the drop-in model on one batch of synthetic token ids:
forward (mask, embedding, 6 relative-attention layers, vocabulary projection), label-smoothed
CE loss, backward, data-parallel gradient all-reduce (N > 1), fused Adam with the Noam rate.
Default workload = BASELINE.json configs[1] ("B"): V=390, 6 layers, d512 (8 heads), L=2048,
batch 16 per GPU, bf16 operands, dropout 0.2 (reference default).

Prints ONE JSON line (rank 0).  `value` = device-timed tokens/s with the ids resident in HBM;
`e2e` = the same step through the public module API with pinned-host ids copied in and the loss
read back every step (each step's loss goes to a pinned slot asynchronously and is read on the host
while the next step runs -- utils.ScalarReadback; the last one before the clock stops).  `--impl reference` times the reference's own
modules (staged by oracle/make_ref.py; the oracle port only if they are absent) on the host cores: the same
optimizer step on a bounded sample of the workload.  At N = 1 the line also carries `cpu_baseline` (that step
timed in the same run) and `gpu_eager_baseline` (the reference modules in eager PyTorch on the same B200).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly the JSON line(s) of this script: everything else a library prints to fd 1 (NCCL's
# version banner when NCCL_DEBUG is set in the environment, ...) is sent to stderr by pointing fd 1 at fd 2
# for the run; emit() writes to the saved stdout.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_STDOUT_FD, (json.dumps(obj) + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (d, V, pad, layers, L, batch per GPU)
    "A": (256, 390, 388, 6, 2048, 2),
    "B": (512, 390, 388, 6, 2048, 16),
    "C": (768, 337, 336, 12, 4096, 4),
}


def flops_per_token(d, V, layers, L):
    """fwd+bwd algorithmic FLOPs per token: layers*(30 d^2 + 9 L d) + 6 d V."""
    return layers * (30 * d * d + 9 * L * d) + 6 * d * V


def traffic_of(kernel: str):
    """(bytes per launch, source file) of the last ncu --set full capture summarised in profiles/roofline_traffic.json."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)[kernel]
        return float(t["bytes"]), t["source"]
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j["bf16_tflops_sustained"],
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def workload_config(cfg_name, Bg, world, dropout):
    """`config` of the JSON line: shared by the GPU arm and the reference arm (the driver compares them)."""
    d, V, pad, layers, L, _ = CONFIGS[cfg_name]
    return {"workload": f"MusicTransformer config {cfg_name}: V={V} {layers}L d{d} h{d // 64} "
                        f"L={L} batch {Bg}/GPU, dropout {dropout}, fwd+loss+bwd+"
                        f"{'allreduce+' if world > 1 else ''}Adam",
            "global_batch": Bg * world, "seq_len": L, "parallelism": f"dp{world}",
            "l2": "per-step activations (several GB) exceed the 126 MB L2; no flush needed"}


def reference_train_step(cfg_name, device, batch, dropout):
    """(step_fn, kind, tokens_per_step): one optimizer step of the reference's own train loop body
    (MT/train.py:258-277 at accum_grad 1: forward in train mode, SmoothCrossEntropyLoss, backward,
    CustomSchedule.step() = Noam rate + torch.optim.Adam.step) on `batch` synthetic sequences.
    kind "reference": the UNMODIFIED MT modules (sources, or the compiled modules oracle/make_ref.py staged);
    kind "port": oracle/restate.py with torch.optim.Adam, only when neither is present."""
    import torch
    from oracle import ref_import
    from oracle import restate as O
    d, V, pad, layers, L, _ = CONFIGS[cfg_name]
    x, y = O.synthetic_ids(batch, L, pad)
    x, y = x.to(device), y.to(device)
    if ref_import.reference_available():
        R = ref_import.load_reference()
        R.config.pad_token = pad
        torch.manual_seed(0)
        m = R.network.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L,
                                       dropout=dropout).to(device)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=0, betas=(0.9, 0.98), eps=1e-9)      # MT/train.py:143
        sched = R.criterion.CustomSchedule(d, optimizer=opt)                             # MT/train.py:151
        crit = R.criterion.SmoothCrossEntropyLoss(0.1, V, pad)                           # MT/train.py:134

        def step():
            sched.optimizer.zero_grad()
            loss = crit(m(x), y)
            loss.backward()
            sched.step()
            return loss
        return step, "reference", batch * L
    p = {k: v.to(device).requires_grad_(True) for k, v in O.init_params(d, V, layers, L, seed=0).items()}
    opt = torch.optim.Adam(list(p.values()), lr=0, betas=(0.9, 0.98), eps=1e-9)

    def step():
        opt.zero_grad()
        loss = O.smooth_ce(O.model_forward(x, p, L, pad), y, 0.1, V, pad)
        loss.backward()
        for g in opt.param_groups:
            g["lr"] = O.noam_rate(max(1, step.n), d)
        step.n += 1
        opt.step()
        return loss
    step.n = 1
    return step, "port", batch * L


def time_cpu_steps(step, steps, warm, budget_s):
    times = []
    t_begin = time.time()
    for it in range(warm + steps):
        t0 = time.time()
        step()
        dt_ = time.time() - t0
        if it >= warm:
            times.append(dt_)
        if time.time() - t_begin > budget_s and len(times) >= 1:
            break
    return times


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the same train step (forward + loss + backward +
    Adam with the Noam rate, fp32) on all host threads of the box; each step is a bounded sample of the
    workload (one sequence of its shape -- batch 16 x 2048 materialises ~40 GB of L x L temporaries on the
    host and takes minutes per step).  Also times BASELINE.json's configs[0] (config A forward + loss)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, V, pad, layers, L, Bg = CONFIGS[args.config]
    if args.batch:
        Bg = args.batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = 1
    step, kind, tokens = reference_train_step(args.config, torch.device("cpu"), Bs, args.dropout)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # keep the whole arm within a few minutes: stop early on a time budget
    times = time_cpu_steps(step, steps, warm, float(os.environ.get("MT_REF_BUDGET_S", "150")))
    ms = 1e3 * sum(times) / len(times)
    val = tokens / (ms / 1e3)
    what = "unmodified MT/{network,layers,criterion,utils}.py" if kind == "reference" else "oracle/restate.py"
    line = {"impl": "reference", "metric": "train_tokens_per_s", "value": val, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": warm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args.config, Bg, max(1, args.gpus), args.dropout),
            "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": cores, "kind": kind,
                             "sample": f"{len(times)} optimizer steps (fwd+loss+bwd+Adam) on {Bs} sequence x {L} tokens "
                                       f"of the workload per step, {what}, torch {torch.__version__} CPU fp32, "
                                       f"{torch.get_num_threads()} threads"},
            "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    # BASELINE.json configs[0]: config A, batch 2, forward + loss on CPU (the reference's own CPU-runnable case)
    try:
        line["baseline_config_a_fwd_loss"] = config_a_forward_loss()
    except Exception as e:                      # noqa: BLE001 -- an extra figure must not lose the line
        line["baseline_config_a_fwd_loss"] = {"error": str(e)[:200]}
    emit(line)


def config_a_forward_loss(budget_s=30.0):
    import torch
    from oracle import ref_import
    from oracle import restate as O
    d, V, pad, layers, L, B = CONFIGS["A"]
    x, y = O.synthetic_ids(B, L, pad)
    if ref_import.reference_available():
        R = ref_import.load_reference()
        R.config.pad_token = pad
        torch.manual_seed(0)
        m = R.network.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0)
        m.train()
        crit = R.criterion.SmoothCrossEntropyLoss(0.1, V, pad)
        fn, kind = (lambda: crit(m(x), y)), "reference"
    else:
        p = O.init_params(d, V, layers, L, seed=0)
        fn, kind = (lambda: O.smooth_ce(O.model_forward(x, p, L, pad), y, 0.1, V, pad)), "port"
    times, loss = [], None
    t_begin = time.time()
    with torch.no_grad():
        for it in range(4):
            t0 = time.time()
            loss = float(fn())
            times.append(time.time() - t0)
            if time.time() - t_begin > budget_s:
                break
    best = min(times[1:]) if len(times) > 1 else times[0]
    return {"tokens_per_s": B * L / best, "ms": 1e3 * best, "loss": loss, "kind": kind,
            "what": f"config A (V={V} {layers}L d{d} L={L} batch {B}) forward + loss, fp32, {os.cpu_count()} host threads"}


def cpu_baseline_leg(cfg_name, dropout, budget_s=25.0):
    """The same reference step on the box's host cores inside the GPU arm's run (N = 1)."""
    import torch
    d, V, pad, layers, L, Bg = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, tokens = reference_train_step(cfg_name, torch.device("cpu"), 1, dropout)
    times = time_cpu_steps(step, 3, 1, budget_s)
    best = min(times)
    what = "unmodified MT modules" if kind == "reference" else "oracle/restate.py"
    return {"value": tokens / best, "unit": "tokens/s", "cores": cores, "kind": kind,
            "sample": f"{len(times)} optimizer steps after 1 warm-up (fwd+loss+bwd+Adam) on 1 sequence x {L} tokens, "
                      f"{what}, torch CPU fp32, {torch.get_num_threads()} threads"}


def gpu_eager_leg(dev, dropout):
    """Second comparator: the reference modules moved to the SAME B200 and run in eager
    PyTorch (fp32, torch defaults: no TF32 matmul) -- config A as it is, config B at batch 2 (batch 16 needs
    ~40 GB of L x L temporaries per layer set).  One optimizer step per iteration, CUDA-event timed."""
    import torch
    out = {}
    for name, cfg, batch in (("config_A_batch2", "A", 2), ("config_B_batch2", "B", 2)):
        try:
            step, kind, tokens = reference_train_step(cfg, dev, batch, dropout)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = {"ms_per_step": ms, "tokens_per_s": tokens / (ms / 1e3), "kind": kind, "batch": batch,
                         "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        except Exception as e:                  # noqa: BLE001
            out[name] = {"error": str(e)[:200]}
        finally:
            step = None
            torch.cuda.empty_cache()
    out["what"] = ("unmodified reference modules .to(cuda), eager PyTorch fp32, fwd+loss+bwd+Adam per step, "
                   "same B200 as the line's value")
    return out


def extra_legs(dev, dropout, precision):
    """N = 1 only, outside the timed region of the headline: the other single-GPU configurations of BASELINE.json as
    short device-timed measurements, so that they are visible in the driver's record -- config C (REMI: V=337, 12 layers,
    d768, 12 heads, L=4096, batch 4; fwd+loss+bwd+Adam) and the decode leg with 256 sequences on one GPU."""
    import torch
    import musicgeneration_b200 as mtb
    from musicgeneration_b200.optim import FlatAdam
    out = {}
    d, V, pad, layers, L, Bg = CONFIGS["C"]
    mtb.config.pad_token = pad
    torch.manual_seed(0)
    model = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=dropout,
                                 precision=precision).to(dev)
    model.train()
    crit = mtb.SmoothCrossEntropyLoss(0.1, V, pad)
    opt = FlatAdam(model, lr=0.0, betas=(0.9, 0.98), eps=1e-9)
    sched = mtb.CustomSchedule(d, optimizer=opt)
    g = torch.Generator().manual_seed(4321)
    xs = [torch.randint(0, pad, (Bg, L), generator=g, dtype=torch.int32).to(dev) for _ in range(2)]
    ys = [torch.randint(0, pad, (Bg, L), generator=g, dtype=torch.int32).to(dev) for _ in range(2)]

    def step(i):
        opt.zero_grad()
        crit(model(xs[i & 1]), ys[i & 1]).backward()
        sched.step()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    K = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out["config_C"] = {"workload": workload_config("C", Bg, 1, dropout)["workload"], "train_tokens_per_s": Bg * L / (ms / 1e3),
                       "ms_per_step": ms, "steps": K, "warmup": 3,
                       "step_model_tflops": flops_per_token(d, V, layers, L) * Bg * L / (ms / 1e3) / 1e12}
    del model, opt, sched, crit
    torch.cuda.empty_cache()
    # decode, 256 sequences x 2047 events on this one GPU (config B's model; the CUDA-graph path: beyond the persistent
    # kernel's 64 rows)
    d, V, pad, layers, L, _ = CONFIGS["B"]
    mtb.config.pad_token = pad
    model = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L, dropout=0.0,
                                 precision=precision).to(dev)
    model.eval()
    seqs, events = 256, L - 1
    prior = torch.randint(0, pad, (seqs, 1), generator=g, dtype=torch.int64).to(dev)
    with torch.no_grad():
        model.generate(prior, length=8, temperature=1.0, top_k=32)
        torch.cuda.synchronize()
        e0.record()
        model.generate(prior, length=events, temperature=1.0, top_k=32)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["decode_256"] = {"decode_events_per_s": seqs * events / (ms / 1e3), "sequences": seqs, "events_per_sequence": events,
                         "top_k": 32, "ms_total": ms}
    del model
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import musicgeneration_b200 as mtb
    from musicgeneration_b200 import ops
    from musicgeneration_b200.optim import FlatAdam
    from musicgeneration_b200.utils import ScalarReadback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, V, pad, layers, L, Bg = CONFIGS[args.config]
    if args.batch:
        Bg = args.batch
    mtb.config.pad_token = pad
    torch.manual_seed(0)
    model = mtb.MusicTransformer(embedding_dim=d, vocab_size=V, num_layer=layers, max_seq=L,
                                 dropout=args.dropout, precision=args.precision).to(dev)
    model.train()
    crit = mtb.SmoothCrossEntropyLoss(0.1, V, pad)
    opt = FlatAdam(model, lr=0.0, betas=(0.9, 0.98), eps=1e-9)
    sched = mtb.CustomSchedule(d, optimizer=opt)
    g = torch.Generator().manual_seed(1234 + rank)
    nbuf = 4
    host_x = [torch.randint(0, pad, (Bg, L), generator=g, dtype=torch.int32).pin_memory() for _ in range(nbuf)]
    host_y = [torch.randint(0, pad, (Bg, L), generator=g, dtype=torch.int32).pin_memory() for _ in range(nbuf)]
    dev_x = [t.to(dev) for t in host_x]
    dev_y = [t.to(dev) for t in host_y]

    def step(x, y):
        opt.zero_grad()
        logits = model(x)
        loss = crit(logits, y)
        loss.backward()
        sched.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = max(1, args.steps), max(3, args.warmup)
    for i in range(W):
        step(dev_x[i % nbuf], dev_y[i % nbuf])
    barrier()
    # ---- device-timed region (inputs resident in HBM) --------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    time.sleep(0.3)
    # per-launch timing of the dominant kernel group (attention backward) with CUDA events on
    # the launching stream
    attn_ev = []
    orig_bwd, orig_fwd = ops.rga_bwd, ops.rga_fwd
    fwd_ev = []

    def timed(fn, store):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*a, **k)
            e1.record()
            store.append((e0, e1))
        return w

    from musicgeneration_b200 import engine as _eng
    _eng.ops.rga_bwd = timed(orig_bwd, attn_ev)
    _eng.ops.rga_fwd = timed(orig_fwd, fwd_ev)
    calls0 = ops.LAUNCH_CALLS[0]
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(dev_x[i % nbuf], dev_y[i % nbuf])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ops.LAUNCH_CALLS[0] - calls0
    _eng.ops.rga_bwd, _eng.ops.rga_fwd = orig_bwd, orig_fwd
    clocks = sampler.stop() if rank == 0 else None
    bwd_ms = sum(a.elapsed_time(b) for a, b in attn_ev) / max(1, len(attn_ev))
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_ev) / max(1, len(fwd_ev))
    # ---- end-to-end region: pinned host ids -> H2D, loss -> D2H every step -------------------
    barrier()
    t0 = time.perf_counter()
    last = 0.0
    rb = ScalarReadback(depth=2)
    for i in range(K):
        x = host_x[i % nbuf].to(dev, non_blocking=True)
        y = host_y[i % nbuf].to(dev, non_blocking=True)
        rb.push(step(x, y))              # async D2H of this step's loss into a pinned slot
        if len(rb) == 2:
            last = rb.pop()              # host reads step i-1's loss while step i runs
    last = rb.drain()[-1]                # ... and the last step's before the clock stops
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- decode leg (BASELINE config D): KV-cached sampling, sequences sharded over ranks ------
    dec = None
    if not args.no_decode:
        seqs, new_events = args.decode_seqs, min(args.decode_events, L - 1)
        model.eval()
        prior = torch.randint(0, pad, (seqs, 1), generator=g, dtype=torch.int64).to(dev)
        with torch.no_grad():
            model.generate(prior, length=8, temperature=1.0, top_k=32)          # warm-up
            barrier()
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0.record()
            out = model.generate(prior, length=new_events, temperature=1.0, top_k=32)
            d1.record()
            barrier()
        dec_ms = d0.elapsed_time(d1)
        model.train()
        assert out.shape == (seqs, 1 + new_events)
        dec = {"ms": dec_ms, "seqs_per_gpu": seqs, "events": new_events}
    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_s, dec["ms"] if dec else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s = float(t[0]), float(t[1])
        if dec:
            dec["ms"] = float(t[2])
    if rank == 0:
        pk = peaks()
        ms_step = ms_total / K
        tokens = Bg * L * world
        val = tokens / (ms_step / 1e3)
        h = d // 64
        U = L * L * 64                                   # per (b,h): one causal-halved contraction
        bwd_flops = 6 * U * Bg * h                       # algorithmic, per launch (one layer)
        fwd_flops = 3 * U * Bg * h
        ach = bwd_flops / (bwd_ms / 1e3) / 1e12 if bwd_ms > 0 else 0.0
        step_flops = flops_per_token(d, V, layers, L) * Bg * L
        line = {
            "metric": "train_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": workload_config(args.config, Bg, world, args.dropout),
            "clocks": clocks,
            "e2e": {"value": tokens * K / e2e_s, "unit": "tokens/s",
                    "h2d_bytes_per_step": 2 * Bg * L * 4, "d2h_bytes_per_step": 4, "last_loss": last,
                    "loss_read": "every step, pinned + async, consumed one step late"},
            "gpu_launches": launches,
            "gpu_launches_what": "launching C-ABI calls of libmt_b200.so inside the timed region (a lower bound on kernels: an "
                                 "attention backward call is 3 kernels, a split-K weight gradient 2; the ncu launch list "
                                 "profiles/r2f_launches_bench_configB.summary.txt has ~172 mt:: kernels per step)",
            "roofline": {"kernel": "rga_bwd (relative attention backward of one layer, one C-ABI call mt_rga_bwd_stash: delta + "
                                   "dK/dV kernel + dQ/dE kernel, both reading the P tiles the forward kept)", "bound": "tensor",
                         "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": ach / pk["tf_sust"],
                         # dram__bytes_read+write of the kernels of one call from the ncu --set full capture of the
                         # config-B layer shape named in profiles/roofline_traffic.json (written by the capture's
                         # summary step); algorithmic operand bytes (q,k,v,O,dO,dq,dk,dv once) are 0.27 GB -- the rest
                         # is the P stash read by either kernel and operand tiles that miss in L2
                         "traffic": traffic_of("rga_bwd")[0] if (args.config == "B" and Bg == 16) else None,
                         "traffic_source": traffic_of("rga_bwd")[1],
                         "peak_source": pk["src"] + " sustained",
                         "ms_per_launch": bwd_ms, "flops_per_launch": bwd_flops,
                         "fwd_ms_per_launch": fwd_ms,
                         "fwd_achieved": fwd_flops / (fwd_ms / 1e3) / 1e12 if fwd_ms > 0 else 0.0},
            "step_model_tflops": step_flops / (ms_step / 1e3) / 1e12,
            "step_frac_of_peak": step_flops / (ms_step / 1e3) / 1e12 / pk["tf_sust"],
        }
        if dec:
            ev = dec["seqs_per_gpu"] * world * dec["events"] / (dec["ms"] / 1e3)
            # HBM roofline of the decode step: per sequence-step the K and V rows of every
            # layer are read once (mean context = events/2), plus the weights once per step
            es = 2 if args.precision == "bf16" else 4
            kv = layers * 2 * (dec["events"] / 2) * d * es * dec["seqs_per_gpu"]
            wts = (layers * (4 * d * d + d * d) + d * V) * es
            line["decode"] = {"metric": "decode_events_per_s", "value": ev, "unit": "events/s",
                              "sequences": dec["seqs_per_gpu"] * world, "events_per_sequence": dec["events"],
                              "top_k": 32, "ms_total": dec["ms"],
                              "hbm_bytes_per_step": kv + wts,
                              "hbm_frac": (kv + wts) * dec["events"] / (dec["ms"] / 1e3) / 1e9 / pk["hbm"],
                              "scaling": "sequences sharded over ranks, no collective"}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(args.config, args.dropout)
        if world == 1 and not args.no_eager:
            del model, opt, sched
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_leg(dev, args.dropout)
            if args.config == "B" and not args.no_extras:
                try:          # (never at the price of the headline line)
                    line["other_configs"] = extra_legs(dev, args.dropout, args.precision)
                except Exception as e:  # noqa: BLE001
                    line["other_configs"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


SWEEP_L = (512, 1024, 2048, 4096, 8192)
SWEEP_DH = (64, 128)


def run_sweep(args):
    """BASELINE configs[4]: relative attention alone, forward + backward, L = 512 .. 8192, head dim 64 /
    128, 32 768 tokens and d = h*dh = 512 per point.  One JSON line per point.  dh = 64 bf16 runs on the
    tcgen05 kernels; dh = 128 does not occur in the drop-in model (MT/layers.py:219 fixes dh = 64) and is
    served by the fp32-math SIMT kernels.  `--impl reference`: the oracle port of the reference op
    sequence (QE^T, mask, pad/reshape skew, softmax, AV) on the host cores, one (batch, head-set) sample,
    L <= 2048 (the L x L x h fp32 intermediates of larger points do not fit the time budget)."""
    import torch
    pk = peaks()
    if args.impl == "reference":
        import math
        from oracle import restate as O
        if int(os.environ.get("RANK", "0")) != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        for dh in SWEEP_DH:
            for L in SWEEP_L:
                if L > 2048:
                    continue
                h = 512 // dh
                g = torch.Generator().manual_seed(L)
                q, k, v = [torch.randn(1, h, L, dh, generator=g).requires_grad_(True) for _ in range(3)]
                E = torch.randn(L, dh, generator=g).requires_grad_(True)
                ar = torch.arange(L)
                mask = (ar[None, :] > ar[:, None])[None, None]
                best = [1e9, 1e9]
                for it in range(3):
                    t0 = time.time()
                    logits = O.rga_scores(q, k, E, L) + (mask.to(torch.int64) * -1e9).float()
                    o = torch.matmul(torch.softmax(logits, -1), v)
                    t1 = time.time()
                    o.sum().backward()
                    t2 = time.time()
                    if it:
                        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
                U = L * L * dh * h
                emit(({"impl": "reference", "metric": "rga_fwd_bwd_ms", "L": L, "dh": dh, "B": 1, "h": h,
                                  "fwd_ms": 1e3 * best[0], "bwd_ms": 1e3 * best[1],
                                  "tokens_per_s": L / (best[0] + best[1]), "cores": os.cpu_count(),
                                  "fwd_tflops": 3 * U / best[0] / 1e12, "kind": "port"}))
        return
    from musicgeneration_b200 import ops
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    for dh in SWEEP_DH:
        for L in SWEEP_L:
            h, B = 512 // dh, max(1, 32768 // L)
            d = h * dh
            g = torch.Generator().manual_seed(L)
            qkv = torch.randn(B, L, 3, h, dh, generator=g).to(torch.bfloat16).to(dev)
            E = torch.randn(L, dh, generator=g).to(torch.bfloat16).to(dev)
            dO = torch.randn(B, L, h, dh, generator=g).to(torch.bfloat16).to(dev)
            strides, ostr = (L * 3 * d, 3 * d, dh), (L * d, d, dh)
            Od = torch.empty(B, L, h, dh, dtype=torch.bfloat16, device=dev)
            lse = torch.empty(B, h, L, device=dev)
            dqkv = torch.zeros(B, L, 3, h, dh, dtype=torch.bfloat16, device=dev)
            dE = torch.zeros(L, dh, device=dev)
            delta = torch.empty(B, h, L, device=dev)
            iters = 10 if dh == 64 else 3
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
            # the pair a training step runs: the forward keeps its P tiles, the backward reads them (head dim 64)
            stash = ops.rga_stash_new(qkv[:, :, 0], E, Od, B, h, L, dh)
            for it in range(-2, iters):
                e = ev[max(it, 0)]
                e[0].record()
                ops.rga_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, ostr, lse, B, h, L, dh, L, True,
                            stash=stash)
                e[1].record()
                ops.rga_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], strides, E, None, Od, dO, ostr, lse, delta,
                            dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], dE, B, h, L, dh, L, True, stash=stash)
                e[2].record()
            torch.cuda.synchronize()
            fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / iters
            bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / iters
            U = L * L * dh * B * h
            emit(({"metric": "rga_fwd_bwd_ms", "L": L, "dh": dh, "B": B, "h": h,
                              "kernels": "tcgen05, training pair (P stash)" if stash is not None else "simt fp32 math",
                              "fwd_ms": fwd, "bwd_ms": bwd, "tokens_per_s": B * L / ((fwd + bwd) / 1e3),
                              "fwd_tflops": 3 * U / (fwd / 1e3) / 1e12, "bwd_tflops": 6 * U / (bwd / 1e3) / 1e12,
                              "fwd_bwd_frac_of_bf16_sustained": 9 * U / ((fwd + bwd) / 1e3) / 1e12 / pk["tf_sust"]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short config C / decode-256 legs (N = 1 only)")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-CUDA reference comparator (N = 1 only)")
    ap.add_argument("--decode-seqs", type=int, default=32)
    ap.add_argument("--decode-events", type=int, default=2047)
    ap.add_argument("--sweep", action="store_true", help="relative-attention microbench sweep (configs[4])")
    args = ap.parse_args()
    if args.sweep:
        run_sweep(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
