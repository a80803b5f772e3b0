"""Build-container only (needs /root/reference): the UNMODIFIED reference task code (dyros_dynamic_walk.py + its torch
utils) on CPU torch with the simulator stubbed (oracle/ref_harness.py: gym.simulate does nothing), i.e. the Python / ATen
share of one reference env-step without PhysX. Writes profiles/r2_reference_task_cpu.md. (BASELINE.md section 4, 2(a);
VERDICT r1 "missing" 6a. It cannot run on the GPU box, so it is a committed measurement, not a bench.py arm.)"""
import os, sys, time, platform
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import ref_harness as RH

assert RH.reference_available(), "needs /root/reference"
rows = []
for threads in (1, min(8, os.cpu_count() or 1)):
    torch.set_num_threads(threads)
    for N in (64, 4096):
        s = RH.make_reference_task(N, seed=0)
        g = torch.Generator().manual_seed(42)
        acts = [torch.rand(N, 13, generator=g) * 2 - 1 for _ in range(8)]
        for i in range(3):
            RH.reference_step(s, acts[i % 8])
        K = 30 if N == 64 else 10
        t0 = time.perf_counter()
        for i in range(K):
            RH.reference_step(s, acts[i % 8])
        dt = (time.perf_counter() - t0) / K
        rows.append((N, threads, dt * 1e3, N / dt))
        print(N, threads, dt * 1e3, N / dt, flush=True)
out = os.path.join(os.path.dirname(__file__), "..", "profiles", "r2_reference_task_cpu.md")
with open(out, "w") as f:
    f.write("# The reference's own task code on the host (simulator stubbed), build container\n\n")
    f.write(f"`tools/time_reference_task_cpu.py`: unmodified `tasks/dyros_dynamic_walk.py` (pre_physics_step + post_physics_step:\n"
            f"PD / delay / noise / termination / reward / reset / observations in ~1,100 ATen calls per step) on CPU torch "
            f"{torch.__version__}, {platform.processor() or platform.machine()}, {os.cpu_count()} logical CPUs; `gym.simulate` "
            f"and the state refreshes are no-ops (PhysX is absent), so this is the Python / ATen share of a reference env-step only.\n\n")
    f.write("| envs | torch threads | ms per step | env-steps/s (task logic only) |\n|---|---|---|---|\n")
    for N, th, ms, r in rows:
        f.write(f"| {N} | {th} | {ms:.1f} | {r:,.0f} |\n")
    f.write("\nFor scale: the oracle port with the dense fp64 physics (bench.py `cpu_baseline`) does ~19 env-steps/s per core; the\n"
            "CUDA path 29.5 M env-steps/s per B200 including physics.\n")
print("wrote", out)
