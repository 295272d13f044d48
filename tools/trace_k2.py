"""Read the FSK2_TIMING trace of block 0 (see fsk_kernel2.cu) and print per-warp event timelines for one frame."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
import pirip_b200 as pb

ncap = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
nsamp = int(0.6 * bench.FS)
plan = pb.FskPlan(bench.FS, bench.RS, bench.M, P=bench.P, Nsym=bench.NSYM, limits=bench.LIMITS, device=0)
mf = plan.max_frames(nsamp)
dev = torch.device("cuda", 0)
iq = bench.gen_captures(torch, ncap, nsamp, 0, dev)
bits = torch.zeros((ncap, mf * plan.nbits), dtype=torch.uint8, device=dev)
nf = torch.zeros(ncap, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream()
plan.demod_batch_dev(iq.data_ptr(), ncap, nsamp, iq.stride(0), bits.data_ptr(), nf.data_ptr(), mf, stream=ctypes.c_void_p(s.cuda_stream))
torch.cuda.synchronize()
N = 512
buf = np.zeros((32, N), np.int64)
L = pb.lib()
L.pirip_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
rc = L.pirip_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), 32 * N)
assert rc == 0, rc
code = buf >> 48
clk = buf & ((1 << 48) - 1)
t0 = clk[clk > 0].min()
names = {1: "pre", 2: "adv/top-done", 3: "frame-top", 10: "stream-sync-done", 14: "decide-lanes-done", 11: "decide+fft-done", 12: "sums-done", 13: "llr-done",
         20: "osc-chunk-start", 22: "osc-chunk-end", 30: "tim-wait", 31: "tim-go", 32: "tim-chunk-done", 40: "w-loop", 41: "w-osc-passed",
         42: "w-mix-done", 43: "w-mixbar-passed", 44: "w-int-done", 50: "fft-end", 51: "fft-start", 52: "task", 53: "task-end", 45: "INT", 46: "mix"}
NW = int(os.environ.get('PIRIP_FSK_K2_NW', '16'))
for w in (NW - 1, NW - 2, NW // 2, NW - 3, 0, NW // 2 - 1):
    ev = [(int(clk[w, i] - t0), int(code[w, i])) for i in range(N) if buf[w, i] != 0]
    print("warp", w, "events", len(ev))
    prev = None
    out = []
    for t, c in ev[:int(os.environ.get('TRN','64'))]:
        out.append("%s@%d(+%d)" % (names.get(c, str(c)), t, 0 if prev is None else t - prev))
        prev = t
    print("  " + " ".join(out))
