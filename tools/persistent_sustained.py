"""Sustained run of the persistent decode step at the bench geometry (what the step's 31 back-to-back launches see, for seconds
instead of 60 ms): ms per launch per window next to NVML power / SM clock / throttle reasons sampled while it runs.
Usage: python tools/persistent_sustained.py [seconds=4] [flags=17]; SR_LIB_PATH selects an alternative build of the same C ABI."""
import os, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import persistent_bringup as PB
from socioreasoner_b200 import lib
import pynvml

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 17
n_layers, vocab = 36, 151936
torch.manual_seed(0)
layers = PB.make_layers(n_layers)
S = PB.State(8, n_layers, 1024, [900 - 7 * i for i in range(8)], vocab)
P = PB.Persistent(S, layers)
lib.decode_persistent_tuning(flags=flags, ring_stages=0, offsets=None)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM), pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)


print("lib:", lib.LIB_PATH, "flags", flags)
P.reset()
for _ in range(5):
    P.run(0)
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
t_end = time.time() + secs
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w = 0
while time.time() < t_end:
    e0.record()
    for _ in range(100):
        P.run(0)
    e1.record(); torch.cuda.synchronize()
    t = time.time()
    recent = [s for s in samples if s[0] > t - 0.2]
    pw = sum(s[1] for s in recent) / max(len(recent), 1)
    ck = sum(s[2] for s in recent) / max(len(recent), 1)
    rs = 0
    for s in recent:
        rs |= s[4]
    print(f"window {w:3d}: {e0.elapsed_time(e1) / 100:.4f} ms/launch  power {pw:6.1f} W  sm {ck:6.0f} MHz  mem {recent[-1][3] if recent else 0} MHz  reasons 0x{rs:x}", flush=True)
    w += 1
stop = True; th.join()
