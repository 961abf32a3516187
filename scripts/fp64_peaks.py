"""FP64 roofline denominators on this box (MEASURED_PEAKS.json holds only HBM and bf16 numbers):
  * cuBLAS DGEMM 8192^3 through torch.matmul(float64) -- best of 10 and a 3 s back-to-back loop (a LIBRARY number, for
    reference only: no library GEMM is on the (T) path);
  * the library's own DMMA.8x8x4 / DFMA issue-rate microbenchmarks (mpqc_t_microbench), which bench.py uses as `peak`;
  * nominal 148 SM x 64 FMA/clk x 2 x clocks.max.sm.
Writes gpurun_out/r02_fp64_peaks.json."""
import ctypes as C, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpqc_b200 import lib as L

lib = L.load()
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3):
    torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps, t0 = 0, time.perf_counter()
e0.record()
while time.perf_counter() - t0 < 3.0:
    torch.matmul(a, b, out=c); reps += 1
    if reps % 8 == 0:
        torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sustained = e0.elapsed_time(e1) * 1e-3 / reps
out = {"cublas_dgemm_8192_tflops_burst": 2.0 * n ** 3 / best * 1e-12,
       "cublas_dgemm_8192_tflops_sustained_3s": 2.0 * n ** 3 / sustained * 1e-12}
for which, name in ((0, "dmma_issue_rate_tflops"), (1, "dfma_issue_rate_tflops"), (2, "dmma_8_warps_per_sm_tflops")):
    tf = C.c_double()
    L.check(lib.mpqc_t_microbench(0, which, C.byref(tf)), "microbench")
    out[name] = tf.value
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm,clocks.sm", "--format=csv,noheader,nounits", "-i", "0"],
                       capture_output=True, text=True).stdout.strip().split(",")
    out["gpu"], out["sm_max_mhz"] = q[0].strip(), float(q[1])
    out["nominal_fp64_tflops"] = 148 * 64 * 2 * float(q[1]) * 1e6 * 1e-12
except Exception:
    pass
out["torch"] = torch.__version__
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_fp64_peaks.json"), "w"), indent=1)
