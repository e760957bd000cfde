#!/usr/bin/env python
"""Benchmark of the PNAPCSAFT hot path (BASELINE.json metric: molecular graphs/s).

Workload at N GPUs (BASELINE config 2 / config 4): the default gnnepcsaft PNA model, fp32 training
step (device collate + forward + MAPE + backward + flat gradient all-reduce for N>1 + fused
AdamW-amsgrad / cosine-restart schedule) on 4096-graph synthetic OGB-featurised batches per GPU
(weak scaling: 4096 graphs per rank per step).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --gpus N --steps K ...  # the reference arithmetic (CPU oracle port)

Prints ONE JSON line on rank 0.
  value     graphs/s of the store-resident path: the dataset lives in HBM in compact dtypes, every step gathers a new
            random 4096-graph batch on the device (gnn_compact_collate) and replays the captured CUDA graph of the step
  e2e       the same step fed from pinned HOST buffers (compact batches, H2D of every batch + D2H of every loss inside
            the timed region)
  pyg_layout  the eager path for batches that arrive as PyG int64 tensors (round-1 headline; ~125 C-ABI calls per
            step) - the per-kernel CUDA-event timings (roofline, roofline_by_kernel, kernel_shares) are taken here,
            because events cannot be recorded around single kernels of a replayed graph
  config1 / config3 / config5 / batch512   the other BASELINE.json configs and the reference's own batch size
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "gnn-epc-saft_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

GRAPHS_PER_GPU = 4096
NUM_BATCHES = 4          # distinct batches cycled through, so nothing is reused across steps
METRIC = "molecular graphs/sec (train step: fwd+bwd+AdamW, fp32)"

# kernels launched per C-ABI call (lower bounds where the count depends on sizes)
KERNELS_PER_CALL = {"gnn_build_csr": 11, "gnn_bn_stats": 2, "gnn_bn_act_bwd": 3, "gnn_bond_table_bwd": 3,
                    "gnn_pna_aggregate_bwd": 2, "gnn_tc_linear_fwd": 2, "gnn_tc_linear_dgrad": 2,
                    "gnn_tc_pna_post_fwd": 2, "gnn_tc_pna_post_dgrad": 2, "gnn_tc_pna_post_wgrad": 2,
                    "gnn_tc_pna_proj_fwd": 2, "gnn_tc_pna_dx_dgrad": 2, "gnn_pna_aggregate_src_bwd": 2,
                    "gnn_tc_pna_postf_fwd": 2, "gnn_tc_pna_postf_wgrad": 2, "gnn_tc_pna_projsrc_fwd": 2,
                    "gnn_tc_pna_dxf_dgrad": 2}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled during the timed regions.  In-process NVML queries (nvidia_ml_py) every
    10 ms; spawning nvidia-smi instead takes ~100 ms per sample and its driver calls perturb the run it measures
    (the end-to-end leg read 5-10 % slow with it), so that is only the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.samples, self.stop_flag, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self.run, daemon=True)
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # the NVML index follows CUDA_VISIBLE_DEVICES only through the UUID / PCI bus id of the torch device
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
            self.nvml = pynvml
            self.max_sm = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append([str(sm), str(self.max_sm)] + ["Active" if mask & self.BITS[k] else "Not Active" for k in self.NAMES])

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.01 if self.nvml is not None else 0.2)

    def start(self):
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        self.thread.join(timeout=6)
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": int(statistics.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_problem(seed, device=None):
    from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
    from gnnepcsaft_b200.train.utils import default_config
    packed = make_packed_graphs(GRAPHS_PER_GPU * NUM_BATCHES, seed=seed)
    cfg = default_config()
    cfg.model = "PNA"
    batches = [batch_from_packed(packed, torch.arange(i * GRAPHS_PER_GPU, (i + 1) * GRAPHS_PER_GPU))
               for i in range(NUM_BATCHES)]
    return packed, cfg, batches


WORKLOAD = ("default gnnepcsaft PNA model (H=64, 6 PNAConv, towers 2, 537829 params), fp32 train step "
            "(collate+fwd+MAPE+bwd+AdamW-amsgrad), 4096 synthetic OGB-featurised graphs (~25 atoms) per GPU per step")


def cpu_reference_step_rate(cfg, packed, sample_graphs, steps, warmup, train=True):
    """The reference arithmetic (oracle/pna_oracle.py = PyG op sequence in plain torch) on the host cores:
    train step (fwd + MAPE + bwd + AdamW amsgrad) or eval forward on `sample_graphs` graphs."""
    from gnnepcsaft_b200.data.synth import batch_from_packed
    from oracle import pna_oracle as O
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = O.create_model(cfg, packed.in_degree_histogram())
    model.train(train)
    opt = O.adamw_amsgrad_reference(model)
    batch = batch_from_packed(packed, torch.arange(sample_graphs))
    target = batch.para.view(-1, cfg.num_para)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            opt.zero_grad()
            loss = O.mape(model(batch), target)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                model(batch)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sample_graphs / statistics.median(times), sum(times), torch.get_num_threads()


def run_reference(args, rank):
    """--impl reference: the reference arithmetic on the host cores, SAME workload (4096-graph train steps)."""
    if rank != 0:
        return
    packed, cfg, batches = build_problem(seed=1)
    sample = GRAPHS_PER_GPU
    torch.set_num_threads(os.cpu_count())
    from oracle import pna_oracle as O
    torch.manual_seed(0)
    model = O.create_model(cfg, packed.in_degree_histogram())
    model.train()
    opt = O.adamw_amsgrad_reference(model)

    def step(i):
        b = batches[i % NUM_BATCHES]
        opt.zero_grad()
        loss = O.mape(model(b), b.para.view(-1, cfg.num_para))
        loss.backward()
        opt.step()
        return float(loss)

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cores = torch.get_num_threads()
    line = {"metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "graphs_per_gpu": GRAPHS_PER_GPU,
                       "note": "CPU arm: one process, all host cores, the same 4096-graph batches"},
            "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} train steps of {sample} graphs (plain-torch restatement of "
                                       "the PyG op sequence; torch_geometric is not installable here)"},
            "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else a library prints (e.g. NCCL's version
    banner) was redirected to stderr by _quiet_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


# Algorithmic (minimum) DRAM bytes per node of the hot kernels at hidden H, towers T, per-tower width F = H,
# F_out = H / T: tensors read + written once, weights / tables ignored.  e = CSR slots per node.
def kernel_bytes_per_node(name, H, T, e):
    F = H
    A = 16 * T * F                      # aggregates [T, 4, F] fp32
    table = {
        "gnn_pna_aggregate_src_fwd": 4 * H + A + 5 * e + 4,          # SURVEY 8d numerator (x row, A out, indices)
        "gnn_pna_aggregate_src_bwd": A + 4 * T * F + 2 * 4 * T * F + 4 * T * F + 5 * e,   # gA, std, Psrc + tab, dPsrc
        "gnn_tc_pna_postf_fwd": 4 * H + A + 8 + 4 * H,               # x, A, scalers -> O
        "gnn_tc_pna_postf_agg_fwd": 4 * H + 4 * T * F + 5 * e + 8 + 4 * H,   # x, Psrc, indices, scalers -> O
        "gnn_tc_pna_postf_wgrad": 4 * H + 4 * H + A + 8,             # dO, x, A, scalers
        "gnn_tc_pna_post_dgrad": 4 * H + 8 + A,                      # dO, scalers -> gA
        "gnn_tc_pna_dxf_dgrad": 4 * T * F + 4 * H + 8 + 2 * 4 * H,   # dPsrc, dO, scalers, dx read + write
        "gnn_tc_pna_projsrc_fwd": 4 * H + 4 * T * F,
        "gnn_tc_pna_projsrc_wgrad": 4 * T * F + 4 * H,
        "gnn_bn_act_bwd": 3 * 4 * H,
        "gnn_bn_act_fwd": 3 * 4 * H,
        "gnn_bn_stats": 4 * H,
    }
    return table.get(name)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config1/3/5 and batch512 legs")
    ap.add_argument("--skip", default="", help="comma-separated extra legs to skip: config1,batch512,config3,config5")
    ap.add_argument("--shares", action="store_true", help="also print the per-kernel time-share table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from gnnepcsaft_b200.data.compact import CompactGraphStore, bucket
    from gnnepcsaft_b200.data.loader import DevicePrefetcher
    from gnnepcsaft_b200.data.synth import batch_from_packed, make_packed_graphs
    from gnnepcsaft_b200.train import models as M
    from gnnepcsaft_b200.train.engine import TrainEngine
    from gnnepcsaft_b200.train.utils import create_model, wide_config

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    packed, cfg, host_batches = build_problem(seed=1 + 1000 * rank)
    torch.manual_seed(0)
    model = create_model(cfg, packed.in_degree_histogram()).to(dev)
    engine = TrainEngine.from_config(model, cfg)
    dev_batches = [b.to(dev) for b in host_batches]
    n_nodes = int(host_batches[0].x.shape[0])
    n_edges = int(host_batches[0].edge_index.shape[1])
    store = CompactGraphStore(packed, dev, self_loops=cfg.add_self_loops)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, count):
        """fn(i) for i in range(count) between two barriers; device time, max over ranks (ms)."""
        sync_all()
        e0.record()
        for i in range(count):
            fn(i)
        e1.record()
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- store-resident training (`value`): new random batch every step, one graph replay per step
    gen = torch.Generator().manual_seed(1234 + rank)

    def plan(count, graphs_per_batch, st):
        out = []
        for _ in range(count):
            ids = torch.randperm(st.num_graphs, generator=gen)[:graphs_per_batch]
            out.append((ids, ids.to(dev)))
        return out

    def precapture(eng, st, batches, fn):
        """Warm-up: run the planned batches' shape buckets until their CUDA graphs exist (real steps, untimed).  Every
        rank runs the SAME number of steps (max over ranks) - a training step contains a collective."""
        todo, seen = [], {}
        for ids, ids_dev in batches:
            key = bucket(*st.sizes(ids))
            if seen.get(key, 0) < 2:
                todo.append((ids, ids_dev))
                seen[key] = seen.get(key, 0) + 1
        count = len(todo)
        if world > 1:
            t = torch.tensor([count], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            count = int(t.item())
        for i in range(count):
            fn(eng, st, *todo[i % len(todo)])
        return len(seen)

    train_fn = lambda eng, st, ids, ids_dev: eng.step_from_store(st, ids, ids_device=ids_dev)     # noqa: E731
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                      # samples clocks from the warm-up through the timed regions
    planned = plan(args.steps, GRAPHS_PER_GPU, store)
    for ids, ids_dev in plan(args.warmup, GRAPHS_PER_GPU, store):
        train_fn(engine, store, ids, ids_dev)
    n_buckets = precapture(engine, store, planned, train_fn)
    replays0 = engine.graph_replays
    losses = []
    ms = timed(lambda i: losses.append(train_fn(engine, store, *planned[i])), args.steps)
    graph_launches = engine.graph_replays - replays0
    value = world * GRAPHS_PER_GPU * args.steps / (ms * 1e-3)
    final_loss = float(losses[-1].item())
    pad_n, pad_s = bucket(*store.sizes(planned[0][0]))

    # ---------------- eager path for PyG-layout int64 batches (round-1 headline) + per-kernel timings
    for i in range(3):
        engine.step(dev_batches[i % NUM_BATCHES])
    folded = model._fold_dst(model._matmul_mode())
    agg_fwd, agg_bwd = (("gnn_pna_aggregate_src_fwd", "gnn_pna_aggregate_src_bwd") if folded else
                        ("gnn_pna_aggregate_fwd", "gnn_pna_aggregate_bwd"))
    prof = M.KernelProfiler(watch=[agg_fwd, agg_bwd])
    M.set_profiler(prof)
    calls0 = M.ABI_CALLS
    ms_pyg = timed(lambda i: engine.step(dev_batches[i % NUM_BATCHES]), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    M.set_profiler(None)
    calls_per_step = (M.ABI_CALLS - calls0) / args.steps
    agg = prof.summary()
    pyg_layout = {"value": world * GRAPHS_PER_GPU * args.steps / (ms_pyg * 1e-3), "unit": "graphs/s",
                  "ms_per_step": ms_pyg / args.steps, "abi_calls_per_step": calls_per_step,
                  "what": "eager launches on device-resident PyG-layout int64 batches (gnn_build_csr per step)"}

    # ---------------- roofline of the fused message+aggregate kernel (timed inside the eager region above)
    peak, peak_kind = peaks()
    H, T, F = cfg.hidden_dim, 2, cfg.hidden_dim
    ep = n_edges + n_nodes                       # self-loops are implicit CSR entries
    alg_bytes = 4 * n_nodes * H + 16 * n_nodes * T * F + 5 * ep + 4 * (n_nodes + 1)
    tot_ms, cnt = agg.get(agg_fwd, (0.0, 0))
    agg_ms = tot_ms / max(cnt, 1)
    achieved = alg_bytes / (agg_ms * 1e-3) / 1e9 if agg_ms > 0 else 0.0
    roofline = {"kernel": "pna_aggregate_fwd_pipe_kernel<1, %s, 3, 3>" % ("true" if folded else "false"),
                "abi_call": agg_fwd,
                "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_kind": peak_kind,
                "frac_of_8000_nominal": achieved / 8000.0,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": agg_ms, "launches_timed": cnt,
                "timed_in": "the eager pyg_layout region of this run (CUDA events around each launch, K steps)"}
    # DRAM bytes per launch of the same kernel on the same workload from the committed ncu --set full capture
    try:
        tname = "r1_agg_src_fwd_traffic.json" if folded else os.path.join("history", "r1_agg_fwd_traffic.json")
        with open(os.path.join(ROOT, "profiles", tname)) as fh:
            tr = json.load(fh)
        if tr.get("workload_nodes") == int(n_nodes):
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_source"] = f"profiles/{tname} (ncu --set full, per launch)"
    except (OSError, ValueError, KeyError):
        pass
    btot, bcnt = agg.get(agg_bwd, (0.0, 0))
    roofline["bwd_avg_launch_ms"] = btot / max(bcnt, 1)

    # ---------------- per-kernel share table (separate instrumented pass, not part of any reported time)
    prof_all = M.KernelProfiler()
    M.set_profiler(prof_all)
    for i in range(2):
        engine.step(dev_batches[i % NUM_BATCHES])
    M.set_profiler(None)
    summ = prof_all.summary()
    total = sum(t for t, _ in summ.values())
    launches_per_step = sum(KERNELS_PER_CALL.get(n, 1) * c for n, (_, c) in summ.items()) / 2
    ranked = sorted(summ.items(), key=lambda kv: -kv[1][0])
    shares = {n: round(t / total, 4) for n, (t, _) in ranked[:8]}
    roofline_by_kernel = []
    for n, (t, c) in ranked[:10]:
        bpn = kernel_bytes_per_node(n, H, T, ep / n_nodes)
        if bpn is None:
            continue
        us = 1e3 * t / c
        gbs = bpn * n_nodes / (us * 1e-6) / 1e9
        roofline_by_kernel.append({"abi_call": n, "share_of_step": round(t / total, 4), "us_per_call": round(us, 1),
                                   "algorithmic_bytes": int(bpn * n_nodes), "achieved_gbs": round(gbs, 1),
                                   "frac_measured_peak": round(gbs / peak, 3), "frac_8000_nominal": round(gbs / 8000.0, 3)})
    if args.shares and rank == 0:
        for n, (t, c) in ranked:
            sys.stderr.write(f"{n:32s} {t / 2:9.3f} ms/step {c // 2:4d} calls/step {100 * t / total:5.1f}%\n")

    # ---------------- end to end from pinned host buffers (`e2e`)
    # Compact batches assembled on the host in pinned memory; TrainEngine.steps_from_host copies each step's batch on a
    # copy stream into preallocated staging buffers while the previous step computes; the step copies it into the
    # bucket's static buffers (device to device) and replays the captured graph; every step's loss is read back to the
    # host (one step late, so the read does not drain the queue).  All copies and reads of the K timed steps happen
    # inside the timed region.
    host_compact = [store.host_batch(ids) for ids, _ in planned[:min(args.steps, 8)]]
    h2d_bytes = host_compact[0].nbytes()
    loss_slots = torch.empty(max(args.steps, 3), dtype=torch.float32, pin_memory=True)   # pinned once, outside the loop

    def host_feed(count):
        for i in range(count):
            yield host_compact[i % len(host_compact)]

    def e2e_run(count):
        out, pending = [], None
        for k, loss_dev in enumerate(engine.steps_from_host(store, host_feed(count), depth=2)):
            host_loss = loss_slots[k:k + 1]
            host_loss.copy_(loss_dev, non_blocking=True)                # D2H of this step's loss
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                out.append(float(pending[0]))
            pending = (host_loss, ev)
        pending[1].synchronize()
        out.append(float(pending[0]))
        return out

    e2e_run(max(3, len(host_compact)))
    e2e_losses = []
    ms_e2e = timed(lambda i: e2e_losses.extend(e2e_run(args.steps)), 1)
    assert len(e2e_losses) == args.steps
    e2e_value = world * GRAPHS_PER_GPU * args.steps / (ms_e2e * 1e-3)

    # ---------------- batched inference (forward only, eval-mode BatchNorm): reported beside the training metric
    infer_fn = lambda eng, st, ids, ids_dev: eng.predict_from_store(st, ids, ids_device=ids_dev)    # noqa: E731
    precapture(engine, store, planned, infer_fn)
    ms_inf = timed(lambda i: infer_fn(engine, store, *planned[i]), args.steps)
    for i in range(2):
        engine.predict(dev_batches[i])
    ms_inf_pyg = timed(lambda i: engine.predict(dev_batches[i % NUM_BATCHES]), args.steps)
    inference = {"value": world * GRAPHS_PER_GPU * args.steps / (ms_inf * 1e-3), "unit": "graphs/s",
                 "ms_per_batch": ms_inf / args.steps,
                 "what": "eval forward of 4096-graph batches gathered from the device store (collate + 6 PNAConv + pool "
                         "+ head; BatchNorm folded into the GEMMs; aggregates never reach HBM), one graph replay each",
                 "pyg_layout": {"value": world * GRAPHS_PER_GPU * args.steps / (ms_inf_pyg * 1e-3), "unit": "graphs/s",
                                "ms_per_batch": ms_inf_pyg / args.steps}}

    # opt-in tile kernel (whole conv stack of a <= 128-node tile per CTA), measured beside the default per-layer path
    if store.tileable:
        M.TILE_INFER = True
        tile_eng = TrainEngine.from_config(model, cfg, distributed=False)
        precapture(tile_eng, store, planned, infer_fn)
        ms_tile = timed(lambda i: infer_fn(tile_eng, store, *planned[i]), args.steps)
        M.TILE_INFER = False
        inference["tile_kernel_opt_in"] = {"value": world * GRAPHS_PER_GPU * args.steps / (ms_tile * 1e-3), "unit": "graphs/s",
                                           "ms_per_batch": ms_tile / args.steps}
        del tile_eng

    extra = {}
    skip = set(args.skip.split(",")) | ({"config1", "batch512", "bf16", "config3", "config5"} if args.no_extra else set())
    if "config1" not in skip:
        # ---------------- BASELINE config 1: default model, eval forward on a 1024-graph batch
        p1 = plan(args.steps, 1024, store)
        precapture(engine, store, p1, infer_fn)
        ms_c1 = timed(lambda i: infer_fn(engine, store, *p1[i]), args.steps)
        b1024 = batch_from_packed(packed, torch.arange(1024)).to(dev)
        for _ in range(2):
            engine.predict(b1024)
        ms_c1_pyg = timed(lambda i: engine.predict(b1024), args.steps)
        extra["config1"] = {"what": "default model, eval forward, 1024-graph batch", "unit": "graphs/s",
                            "value": world * 1024 * args.steps / (ms_c1 * 1e-3), "ms_per_batch": ms_c1 / args.steps,
                            "pyg_layout": {"value": world * 1024 * args.steps / (ms_c1_pyg * 1e-3),
                                           "ms_per_batch": ms_c1_pyg / args.steps}}

    if "batch512" not in skip:
        # ---------------- the reference's own batch size (configs/default.py:22): 512-graph training steps
        p512 = plan(args.steps, 512, store)
        precapture(engine, store, p512, train_fn)
        ms_512 = timed(lambda i: train_fn(engine, store, *p512[i]), args.steps)
        b512 = [batch_from_packed(packed, torch.arange(i * 512, (i + 1) * 512)).to(dev) for i in range(4)]
        for i in range(2):
            engine.step(b512[i])
        ms_512_pyg = timed(lambda i: engine.step(b512[i % 4]), args.steps)
        extra["batch512"] = {"what": "default model, fp32 train step, 512 graphs per GPU per step", "unit": "graphs/s",
                             "value": world * 512 * args.steps / (ms_512 * 1e-3), "ms_per_step": ms_512 / args.steps,
                             "pyg_layout": {"value": world * 512 * args.steps / (ms_512_pyg * 1e-3),
                                            "ms_per_step": ms_512_pyg / args.steps}}

    if "bf16" not in skip:
        # ---------------- the same 4096-graph workload with the dense contractions as bf16-in / fp32-accumulate tiles
        # (north_star item 3); everything else (aggregation, BatchNorm, pooling, optimizer) stays fp32.  The headline
        # stays the fp32-parity (TF32x3) path; the bf16 tolerances are the measured ones in tests/test_model_gpu.py.
        torch.manual_seed(0)
        bmodel = create_model(cfg, packed.in_degree_histogram()).to(dev)
        bmodel.matmul_precision = "bf16"
        beng = TrainEngine.from_config(bmodel, cfg)
        precapture(beng, store, planned, train_fn)
        ms_b = timed(lambda i: train_fn(beng, store, *planned[i]), args.steps)
        precapture(beng, store, planned, infer_fn)
        ms_bi = timed(lambda i: infer_fn(beng, store, *planned[i]), args.steps)
        extra["bf16_gemms"] = {"what": "default model, 4096 graphs per GPU per step, dense contractions with bf16 operands "
                                       "and fp32 accumulation (matmul_precision='bf16'); the fused aggregate+post kernel and the "
                                       "tensor-map operand path exist for TF32x3 only, so at hidden 64 this mode trains "
                                       "slower than the fp32-parity headline and only its eval forward is faster",
                               "unit": "graphs/s",
                               "train": world * GRAPHS_PER_GPU * args.steps / (ms_b * 1e-3),
                               "train_ms_per_step": ms_b / args.steps,
                               "eval_forward": world * GRAPHS_PER_GPU * args.steps / (ms_bi * 1e-3),
                               "eval_ms_per_batch": ms_bi / args.steps}
        del beng, bmodel

    if "config3" not in skip:
        # ---------------- BASELINE config 3: screening of a 10 M-graph stream, sharded over the ranks by node count
        # (strong scaling, no collective).  Every rank holds the same 32768-graph store in HBM (compact dtypes); the
        # stream is that store passed `passes` times; rank r takes the r-th contiguous node-balanced range.
        sc_packed = make_packed_graphs(32768, seed=77)
        sc_store = CompactGraphStore(sc_packed, dev, self_loops=cfg.add_self_loops, keep_host_mirror=False)
        passes = -(-10_000_000 // sc_store.num_graphs)
        bounds = sc_store.shard_bounds(world)
        lo, hi = bounds[rank]
        engine.screen(sc_store, lo, hi, max_nodes=1 << 19, out=False, passes=1)       # warm-up + capture
        engine.screen(sc_store, lo, hi, max_nodes=1 << 19, out=False, passes=1)
        info = {}

        def run_screen(_):
            _, nmb, nodes = engine.screen(sc_store, lo, hi, max_nodes=1 << 19, out=False, passes=passes)
            info["micro_batches"], info["nodes"] = nmb, nodes

        ms_sc = timed(run_screen, 1)
        shard_nodes = torch.tensor([float(sc_store.node_count[a:b].sum()) for a, b in bounds])
        extra["config3"] = {"what": f"eval forward over a stream of {passes * sc_store.num_graphs} graphs "
                                    f"({sc_store.num_graphs} distinct synthetic graphs resident in HBM, {passes} passes), "
                                    "contiguous ranges balanced by node count over the ranks, micro-batches of <= 524288 "
                                    "nodes in one shape bucket (one captured CUDA graph), predictions discarded",
                            "unit": "graphs/s", "scaling": "strong",
                            "value": passes * sc_store.num_graphs / (ms_sc * 1e-3), "seconds": ms_sc * 1e-3,
                            "micro_batches_per_pass_rank0": info["micro_batches"],
                            "graphs_per_micro_batch": round((hi - lo) / info["micro_batches"], 1),
                            "rank_node_imbalance": float(shard_nodes.max() / shard_nodes.mean() - 1.0)}
        del sc_store, sc_packed

    if "config5" not in skip:
        # ---------------- BASELINE config 5: wide / deep variant (hidden 256, 6 conv layers, post_layers 2, no
        # self-loops), skewed graph sizes (5-150 atoms), 1024-graph batches; bf16 node-update GEMMs and TF32x3
        wcfg = wide_config()
        wcfg.model = "PNA"
        wpacked = make_packed_graphs(4096, seed=4 + rank, skewed=True)
        wstore = CompactGraphStore(wpacked, dev, self_loops=wcfg.add_self_loops, keep_host_mirror=False)
        c5 = {"what": "wide model (hidden 256, 6 PNAConv, post_layers 2, no self-loops), skewed 5-150-atom graphs, "
                      "1024 graphs per GPU per step", "unit": "graphs/s"}
        for mode in ("bf16", "tf32x3"):
            torch.manual_seed(0)
            wmodel = create_model(wcfg, wpacked.in_degree_histogram()).to(dev)
            wmodel.matmul_precision = mode
            weng = TrainEngine.from_config(wmodel, wcfg)
            pw = plan(max(args.steps // 2, 4), 1024, wstore)
            for ids, ids_dev in pw[:2]:
                train_fn(weng, wstore, ids, ids_dev)
            precapture(weng, wstore, pw, train_fn)
            ms_w = timed(lambda i: train_fn(weng, wstore, *pw[i]), len(pw))
            precapture(weng, wstore, pw, infer_fn)
            ms_wi = timed(lambda i: infer_fn(weng, wstore, *pw[i]), len(pw))
            c5[mode] = {"train": world * 1024 * len(pw) / (ms_w * 1e-3), "train_ms_per_step": ms_w / len(pw),
                        "eval_forward": world * 1024 * len(pw) / (ms_wi * 1e-3), "eval_ms_per_batch": ms_wi / len(pw)}
            del weng, wmodel
        c5["nodes_per_batch"] = int(wstore.sizes(pw[0][0])[0])
        extra["config5"] = c5
        del wstore, wpacked

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, secs, cores = cpu_reference_step_rate(cfg, packed, GRAPHS_PER_GPU, steps=2, warmup=1)
        cpu_baseline = {"value": v, "unit": "graphs/s", "cores": cores, "kind": "port",
                        "sample": f"2 train steps of {GRAPHS_PER_GPU} graphs ({secs:.1f} s) with the plain-torch "
                                  "restatement of the PyG op sequence (oracle/pna_oracle.py); torch_geometric is not "
                                  "installable"}
        if "config1" in extra:
            v1, secs1, _ = cpu_reference_step_rate(cfg, packed, 1024, steps=3, warmup=1, train=False)
            extra["config1"]["cpu_baseline"] = {"value": v1, "unit": "graphs/s", "cores": cores, "kind": "port",
                                                "sample": f"3 eval forwards of 1024 graphs ({secs1:.1f} s)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "graphs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD,
                           "graphs_per_gpu": GRAPHS_PER_GPU, "nodes_per_batch": n_nodes, "edges_per_batch": n_edges,
                           "padded_nodes": pad_n, "padded_csr_slots": pad_s, "shape_buckets": n_buckets,
                           "feed": "compact store in HBM (uint8 features, int32 CSR), a new random batch per step by "
                                   "gnn_compact_collate, one CUDA-graph replay per step",
                           "parallelism": f"dp{world}", "l2": "activations of one step (~2.4 GB) exceed the 126 MB L2; "
                                                             "every step is a different random batch"},
                "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                        "what": "pinned host compact batches -> H2D on a copy stream into staging buffers -> captured step "
                                "-> D2H loss"},
                "gpu_launches": int(launches_per_step * args.steps),
                "graph_launches": graph_launches, "kernels_per_step": launches_per_step,
                "roofline": roofline, "roofline_by_kernel": roofline_by_kernel, "cpu_baseline": cpu_baseline,
                "clocks": clocks, "kernel_shares": shares, "pyg_layout": pyg_layout,
                "inference": inference, **extra, "final_loss": final_loss}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
