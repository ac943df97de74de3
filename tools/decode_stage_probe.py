"""Where do the decode-stage milliseconds of one bench step go?  Times the stage (CUDA events at the engine's stage
boundaries) with the decode step replayed as a CUDA graph and launched eagerly, and with plain / cooperative launches of
the persistent kernel.  Usage: python tools/decode_stage_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from socioreasoner_b200 import lib
from socioreasoner_b200.config import SocioReasonerConfig
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration
from socioreasoner_b200.rollout import DataParallelRollout

dev = torch.device("cuda", 0)
cfg = SocioReasonerConfig()
model = SocioReasonerB200ForConditionalGeneration.random_init(cfg, seed=42, device=dev)
ro = DataParallelRollout(model, max_new_tokens=bench.NEW_TOKENS)
devb = bench.to_device(bench.build_batch(0, cfg), dev)
eng = model.engine


def stage(tag):
    for _ in range(2):
        ro.step(devb, bench.TILES_PER_GPU); ro.gather(bench.TILES_PER_GPU)
    torch.cuda.synchronize()
    res = []
    for _ in range(3):
        eng.stage_events = []
        ro.step(devb, bench.TILES_PER_GPU); ro.gather(bench.TILES_PER_GPU)
        torch.cuda.synchronize()
        evs, eng.stage_events = eng.stage_events, None
        res.append({evs[i + 1][0]: evs[i][1].elapsed_time(evs[i + 1][1]) for i in range(len(evs) - 1)})
    d = min(r["decode"] for r in res)
    print(f"{tag:44s} vit+prefill {min(r['vit+prefill'] for r in res):7.2f} ms  decode {d:7.2f} ms = {d / (bench.NEW_TOKENS - 1):.3f} ms/token", flush=True)


stage("graph replay")
eng.use_cuda_graph = False
stage("eager launches")
eng.use_cuda_graph = True
eng._graphs.clear()
eng.set_persistent_decode(False)
stage("per-layer kernels, graph replay")
