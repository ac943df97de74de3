"""Why is the persistent decode step slower inside the bench step (2.0-2.1 ms) than stand-alone (1.81 ms)?  Probe: ~70 ms of
back-to-back tcgen05 GEMMs (the prefill's load), then 31 decode launches, each timed with CUDA events, SM clock / power sampled
through NVML the whole time; repeated 6 times like consecutive bench steps.  Prints per-launch times by position and the clock
samples by phase.  Usage: python tools/decode_after_gemm_probe.py"""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib
import pynvml

n_layers, vocab = 36, 151936
torch.manual_seed(0)
layers = PB.make_layers(n_layers)
S = PB.State(8, n_layers, 1024, [900 - 7 * i for i in range(8)], vocab)
P = PB.Persistent(S, layers)
dev = torch.device("cuda")
a = torch.randn((6936, 2048), device=dev).bfloat16()
w = (torch.randn((22016, 2048), device=dev) * 0.02).bfloat16()
out = torch.empty((6936, 11008), dtype=torch.bfloat16, device=dev)
gemm = lambda: lib.gemm(a, w, out=out, epilogue=lib.EPI_SWIGLU)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))


P.reset()
for _ in range(3):
    P.run(0); gemm()
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
N_GEMM, N_DEC, REPS = 160, 31, 6
per_pos = [[] for _ in range(N_DEC)]
marks = []
for rep in range(REPS):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N_DEC + 1)]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    g0.record()
    for _ in range(N_GEMM):
        gemm()
    g1.record()
    evs[0].record()
    for i in range(N_DEC):
        P.run(0)
        evs[i + 1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gemm_ms = g0.elapsed_time(g1)
    dec = [evs[i].elapsed_time(evs[i + 1]) for i in range(N_DEC)]
    marks.append((t0, t0 + gemm_ms * 1e-3, t1))
    for i, d in enumerate(dec):
        per_pos[i].append(d)
    print(f"rep {rep}: GEMM phase {gemm_ms:.1f} ms ({N_GEMM} launches), decode phase {sum(dec):.1f} ms: first {dec[0]:.3f}, 5th {dec[4]:.3f}, 15th {dec[14]:.3f}, last {dec[-1]:.3f} ms", flush=True)
# Same work, INTERLEAVED: 5 GEMM launches between consecutive decode launches (what a rollout that overlaps the prefill of the
# next batch with the decode of the current one would submit).  If the clock governor, not the energy, is what slows the
# sequential arrangement down, this total is lower.
seq_ms, inter_ms = [], []
for rep in range(REPS):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(N_GEMM):
        gemm()
    for i in range(N_DEC):
        P.run(0)
    a1.record(); torch.cuda.synchronize()
    seq_ms.append(a0.elapsed_time(a1))
for rep in range(REPS):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    done = 0
    for i in range(N_DEC):
        n = (N_GEMM * (i + 1)) // N_DEC - done
        for _ in range(n):
            gemm()
        done += n
        P.run(0)
    a1.record(); torch.cuda.synchronize()
    inter_ms.append(a0.elapsed_time(a1))
print("sequential  (160 GEMMs, then 31 decode steps), ms per rep:", " ".join(f"{v:.1f}" for v in seq_ms))
print("interleaved (5 GEMMs between decode steps),    ms per rep:", " ".join(f"{v:.1f}" for v in inter_ms))
stop = True; th.join()
print("decode launch time by position (mean over reps 1..):", " ".join(f"{sum(v[1:]) / len(v[1:]):.3f}" for v in per_pos))
for name, sel in (("GEMM phase", lambda t, m: m[0] <= t < m[1]), ("decode phase", lambda t, m: m[1] <= t < m[2])):
    ck = [c for (t, c, p) in samples for m in marks[1:] if sel(t, m)]
    pw = [p for (t, c, p) in samples for m in marks[1:] if sel(t, m)]
    if ck:
        ck.sort()
        print(f"{name}: {len(ck)} NVML samples, SM clock min {ck[0]} / median {ck[len(ck) // 2]} / max {ck[-1]} MHz, power mean {sum(pw) / len(pw):.0f} W")
