#!/usr/bin/env python
"""Benchmark of the speech-encoder hot path: SLU train step (fwd + bwd + grad all-reduce + Adam) of
experiments/unfreeze_all_layers on synthetic 16 kHz 4 s utterances, batch 256 per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5            # this repo's sm_100a path
    torchrun --nproc-per-node N bench.py --gpus N ...           # one process per GPU, NCCL
    python bench.py --impl reference ...                        # the reference's CPU execution (oracle port)

Prints ONE JSON line (rank 0).  `value` = utterances/s with inputs resident in HBM (device-timed, max
over ranks); `e2e` = the same step through the public API (models.Model.forward) from pinned HOST
buffers with the H2D copy and the D2H loss read inside the timed region.  `roofline` describes the
dominant kernel, `cpu_baseline` the reference-style CPU execution on this box's host cores.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_SAMPLES = 64000            # 4 s @ 16 kHz
WORKLOAD = ("experiments/unfreeze_all_layers.cfg SLU train step (all layers unfrozen, dropout 0.5, Adam), "
            "4 s @16 kHz synthetic utterances")
GRU_T = (400, 200, 100, 50, 25)
GRU_I = (60, 256, 256, 256, 256)


def host_threads():
    """Threads for the CPU reference legs: torch's intra-op pool gets SLOWER beyond ~16-32 threads on this op mix
    (measured: 128 threads = 0.18 utt/s vs 8 threads = 12 utt/s), so use the cores it can actually exploit."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get("SLU_REF_THREADS", "16"))))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.lo, self.hi = [], None, index, 0, None

    def __enter__(self):
        self.lo = len(self.rows)             # rows from here on were sampled inside the timed region
        return self

    def __exit__(self, *a):
        time.sleep(0.12)                     # at least one 100 ms sample lands inside short regions
        self.hi = len(self.rows)

    def start(self):
        """Launch nvidia-smi ahead of the timed region (its start-up takes longer than a short run)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: [self.rows.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)    # gone before anything else is timed
            except Exception:
                self.proc.kill()
            self.thread.join(timeout=2)
            self.proc = None

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.rows[self.lo:self.hi]:
            f = [v.strip() for v in l.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
                reasons |= {n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")}
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


_JSON_FD = None


def _claim_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, warnings) goes to stderr; the one JSON line is written to the
    real stdout by _emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port with the
    reference's execution structure incl. its 80x conv loop), all host threads, bounded sample."""
    if rank != 0:
        return
    import torch
    from oracle import ref_port, torch_ref as R
    cores = host_threads()
    torch.set_num_threads(cores)
    B = args.ref_batch
    steps, warmup = max(1, args.steps), max(0, args.warmup)      # a step = one train step on a bounded sample (ref_batch utterances)
    sec = ref_port.train_steps(R.synthetic_params(seed=0), B, T_SAMPLES, steps, warmup, device="cpu", loop80=True)
    val = B / sec
    line = {"impl": "reference", "metric": "utterances_per_sec_train_step", "value": val, "unit": "utt/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "samples_per_utt": T_SAMPLES, "batch_per_step": B,
                       "note": "bounded sample of the workload: the same train step on batch_per_step utterances per step"},
            "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port",
                             "sample": "%d train steps of batch %d x 4 s (reference execution structure incl. 80x conv loop), "
                                       "torch CPU %d threads" % (steps, B, cores)},
            "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="utterances per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--background-prefetch", action="store_true", help="stage batch i+1 in a helper thread instead of the consumer's")
    ap.add_argument("--launch-detail", action="store_true", help="print every launch of one step with its sizes and device time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-gpu", action="store_true", help="also time the reference-structured port on this GPU (cuDNN)")
    ap.add_argument("--eval-dropout", action="store_true", help="disable dropout (debug)")
    args = ap.parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("end-to-end-slu_b200")
    pkg._lib.load()                                          # fail loudly if the CUDA library is missing
    pkg.dp.install()
    import models
    cfgmod = importlib.import_module("end-to-end-slu_b200.config")
    cfg = cfgmod.read_config(os.path.join(ROOT, "configs", "unfreeze_all_layers.cfg"))
    cfg.pretraining_type = 0                                 # random init, nothing frozen == fully unfrozen
    cfg.Sy_intent, cfg.values_per_slot = cfgmod.fsc_intent_table()
    cfg.num_phonemes = 42
    torch.manual_seed(cfg.seed)
    model = models.Model(cfg)
    model.train()
    if args.eval_dropout:
        model.eval()
    params = [p for p in model.parameters()]
    opt = torch.optim.Adam(params, lr=cfg.training_lr)       # what training.py:19 constructs
    B = args.batch
    NB = 4                                                   # rotating input sets: 4 x 65.5 MB > 126 MB L2
    gen = torch.Generator().manual_seed(1234 + rank)
    xs_host = [(0.1 * torch.randn(B, T_SAMPLES, generator=gen)).pin_memory() for _ in range(NB)]
    ys_host = [torch.stack([torch.randint(0, v, (B,), generator=gen) for v in (6, 14, 4)], 1).pin_memory() for _ in range(NB)]
    xs_dev = [x.cuda(non_blocking=True) for x in xs_host]
    ys_dev = [y.cuda(non_blocking=True) for y in ys_host]

    def step(x, y):
        loss, acc = model(x, y)
        opt.zero_grad()
        loss.backward()
        opt.step()                                           # pre-step hook = the single gradient all-reduce
        return loss, acc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, host):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        if host == "prefetch":                                        # the Trainer loop over a DevicePrefetcher-wrapped loader
            batches = ((xs_host[i % NB], ys_host[i % NB]) for i in range(k))
            for x, y in pkg.loader.DevicePrefetcher(batches, background=args.background_prefetch):   # H2D of batch i+1 under step i
                loss, _ = step(x, y)
                loss.item()                                           # D2H read of the step's result, every step
        for i in range(k if host != "prefetch" else 0):
            if host == "serial":
                loss, _ = step(xs_host[i % NB], ys_host[i % NB])     # H2D inside Model.forward (models.py: x.cuda())
                loss.item()
            else:
                step(xs_dev[i % NB], ys_dev[i % NB])
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    clk = ClockSampler(local).start()
    for i in range(max(args.warmup, 3)):
        step(xs_dev[i % NB], ys_dev[i % NB])
    timed(3, host="prefetch")                                # untimed warm-up of the host-fed paths too (copy stream, its
    timed(3, host="serial")                                  # allocator pool, pinned staging)
    for _ in range(30):                                      # nvidia-smi is up and sampling before the timed region starts
        if clk.rows or clk.proc is None:
            break
        time.sleep(0.1)
    calls0 = pkg._lib.stats["calls"]
    with clk:
        ms_dev = timed(args.steps, host=None)
    launches = pkg._lib.stats["calls"] - calls0
    clk.stop()
    ms_e2e = timed(args.steps, host="prefetch")
    ms_e2e_serial = timed(args.steps, host="serial")
    per_step = ms_dev / args.steps
    value = world * B / (per_step * 1e-3)
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    # ---- per-kernel device time (CUDA events around every C-ABI launch, separate untimed pass) -----
    pkg.ops.OVERLAP = False                                  # serialise the side-stream launches: clean per-kernel durations
    pkg._lib.profile_begin(detail=args.launch_detail)
    for i in range(3):
        step(xs_dev[i % NB], ys_dev[i % NB])
    if args.launch_detail and rank == 0:                     # per-launch table (sizes, ms) of the last profiled step -> stderr
        det = pkg._lib.profile_detail()
        for n, a, ms in det[2 * len(det) // 3:]:
            print("launch %-22s %-60s %8.1f us" % (n, a, ms * 1e3), file=sys.stderr)
    prof = pkg._lib.profile_end()                            # {name: [ms, ...]}
    pkg.ops.OVERLAP = True
    hbm_peak, tf_burst, tf_sust, how = peaks()
    kern = {k: {"launches_per_step": len(v) / 3, "ms_per_step": sum(v) / 3} for k, v in prof.items()}
    gru_names = [k for k in prof if k.startswith("slu_gru_fwd")]
    roofline, extra = None, {}
    if gru_names:
        name = gru_names[0]
        # algorithmic flops of the recurrent contraction h.W_hh^T per launch, summed over the 5 layers / step
        flops = sum(2 * 2 * t * 384 * 128 * B for t in GRU_T)
        sec = kern[name]["ms_per_step"] * 1e-3
        ach = flops / sec / 1e12
        nr_rows = 16 if B >= 1184 else (8 if B >= 592 else 4)         # batch rows per CTA (csrc/gru_tc.cu pick_rows)
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture (layer 0, B=256)
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get(name, {}).get("bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s",
                    "frac": ach / tf_sust, "traffic": traffic, "peak_source": how + " bf16 sustained (kernel timed inside a step)",
                    "traffic_note": "bytes of the layer-0 launch (T=400, the largest of the 5); its algorithmic bytes are 996 MB",
                    "executed_tflops": (32 // nr_rows if nr_rows < 16 else 3) * ach if pkg.ops.GRU_IMPL == "tc" else ach,
                    "note": "persistent-GRU forward, 5 launches/step summed; algorithmic flops = 2 dirs * T_l * 2*384*128 * B over the "
                            "5 layers (h.W_hh only).  The N=16 MMA tile carries %d batch rows, hi and lo stacked along N, times the W_hi / "
                            "W_lo passes: the tensor core executes %dx the algorithmic flops.  The recurrence is a dependency chain "
                            "(%d CTAs of %d batch rows): see DESIGN.md section 4" % (nr_rows, 32 // nr_rows if nr_rows < 16 else 3,
                                                                                    2 * ((B + nr_rows - 1) // nr_rows), nr_rows)}
    sinc_names = [k for k in prof if k.startswith("slu_sincconv_fwd")]
    if sinc_names:
        name = sinc_names[0]
        bytes_ = B * (4 * T_SAMPLES + 4 * 80 * 400)
        sec = kern[name]["ms_per_step"] * 1e-3
        tr_s = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tr_s = json.load(f).get(name, {}).get("bytes_per_launch")
        except Exception:
            pass
        extra["roofline_sincconv"] = {"kernel": name, "bound": "hbm", "achieved": bytes_ / sec / 1e9, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": bytes_ / sec / 1e9 / hbm_peak, "traffic": tr_s,
                                      "note": "algorithmic bytes = B*(4*T + 4*80*L1) = read the waveform once, write the pooled frames once"}

    line = None
    if rank == 0:
        cpu_base = None
        if not args.no_cpu_baseline and world == 1:
            from oracle import ref_port, torch_ref as R
            cores = host_threads()
            torch.set_num_threads(cores)
            sec = ref_port.train_steps(R.synthetic_params(seed=0), args.ref_batch, T_SAMPLES, 2, 1, device="cpu", loop80=True)
            cpu_base = {"value": args.ref_batch / sec, "unit": "utt/s", "cores": cores, "kind": "port",
                        "sample": "2 train steps of batch %d x 4 s, reference execution structure (80x conv loop, nn.GRU), "
                                  "torch CPU %d threads" % (args.ref_batch, cores)}
            if args.ref_gpu:
                sec_g = ref_port.train_steps(R.synthetic_params(seed=0), B, T_SAMPLES, 3, 1, device="cuda", loop80=True)
                extra["ref_port_on_this_gpu"] = {"value": B / sec_g, "unit": "utt/s", "ms_per_step": sec_g * 1e3,
                                                 "what": "reference-structured port (cuDNN GRU, cuDNN conv, 80x conv loop), wall clock"}
        line = {"metric": "utterances_per_sec_train_step", "value": value, "unit": "utt/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tensor-core contractions as bf16 hi/lo 3-pass split, fp32 accumulate)",
                "data": "synthetic", "impl": "ours",
                "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world,
                           "samples_per_utt": T_SAMPLES, "parallelism": "dp%d" % world,
                           "l2": "inputs rotate over 4 batches (262 MB) > 126 MB L2; activations ~1 GB/step"},
                "e2e": {"value": e2e_value, "unit": "utt/s", "h2d_bytes_per_step": B * T_SAMPLES * 4 + B * 3 * 8,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                        "h2d": "pinned host batches, copy of batch i+1 on a copy stream under step i (loader.DevicePrefetcher)",
                        "ms_per_step_serial_copy": ms_e2e_serial / args.steps},
                "gpu_launches": launches, "kernels": kern, "roofline": roofline, "cpu_baseline": cpu_base,
                "clocks": clk.summary(), "allreduce": dict(pkg.dp.stats)}
        line.update(extra)
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
