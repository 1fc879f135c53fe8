"""Developer tool: where does a step of the tcgen05 GRU forward kernel spend its cycles?
Phases (accumulated clock64 deltas for thread 0 [MMA issuer warp] and thread 128): 0 wait MMAs, 1 tcgen05.ld, 2 TMA ring wait,
3 gate math + operand stores, 4 fences, 5 __syncthreads, 6 MMA issue, 7 global stores + loop tail.
Needs a library built with the debug switches: SLU_KERNEL_DEBUG=1 python __graft_entry__.py (the default build has none)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("end-to-end-slu_b200")
lib = pkg._lib.load()
B, T, I = 256, 400, 60
gru = torch.nn.GRU(I, 128, batch_first=True, bidirectional=True).cuda()
x = torch.randn(B, T, I, device="cuda")
names = ["wait_mma", "tmem_ld", "tma_wait", "math+sts", "fences", "syncthreads", "mma_issue", "stores+tail"]
for prec in ("bf16x3", "fp16"):
    pkg.ops.set_gru_precision(prec)
    for train in (False, True):
        buf = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.slu_debug_gru_phase_clocks(buf.data_ptr())
        xx = x.clone().requires_grad_(train)
        y = pkg.ops.bigru(xx, gru, (0.5, 77) if train else None, 2)
        torch.cuda.synchronize()
        lib.slu_debug_gru_phase_clocks(None)
        v = buf.cpu().view(2, 8).double() / T
        print(prec, "stash" if train else "infer", "cycles/step: thread0", {n: int(c) for n, c in zip(names, v[0])}, "sum", int(v[0].sum()))
        print(" " * 20, "thread128", {n: int(c) for n, c in zip(names, v[1])}, "sum", int(v[1].sum()))
