"""Measures the dense TF32 tensor-pipe peak of this GPU (cuBLAS, torch.matmul with allow_tf32) the same way
MEASURED_PEAKS.json measures bf16: 8192^3, best of 10 (burst) and back to back for ~2 s (sustained).  Prints one JSON line."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3): a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0; t0 = time.time(); e0.record()
while time.time() - t0 < 2.0:
    for _ in range(20): a @ b
    reps += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
fl = 2.0 * n ** 3
ab, bb = a.bfloat16(), b.bfloat16()
for _ in range(3): ab @ bb
bestb = 1e9
for _ in range(10):
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(); ab @ bb; e3.record(); torch.cuda.synchronize()
    bestb = min(bestb, e2.elapsed_time(e3))
print(json.dumps({"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl * reps / e0.elapsed_time(e1) / 1e9,
                  "bf16_tflops_here": fl / bestb / 1e9, "how": "torch.matmul fp32 inputs, allow_tf32=True, 8192^3, best of 10 / 2 s back to back"}))
