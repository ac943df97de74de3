"""One bench step under the CUDA profiler range (for ncu --profile-from-start off): warm-up, then exactly one step."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from socioreasoner_b200.config import SocioReasonerConfig
from socioreasoner_b200.modeling import SocioReasonerB200ForConditionalGeneration
from socioreasoner_b200.rollout import DataParallelRollout

graph = os.environ.get("SR_PROFILE_GRAPH", "1") == "1"
dev = torch.device("cuda", 0)
cfg = SocioReasonerConfig()
model = SocioReasonerB200ForConditionalGeneration.random_init(cfg, seed=42, device=dev)
model.engine.use_cuda_graph = graph
ro = DataParallelRollout(model, max_new_tokens=bench.NEW_TOKENS)
host = bench.build_batch(0, cfg)
devb = bench.to_device(host, dev)
for _ in range(2):
    ro.step(devb, bench.TILES_PER_GPU); ro.gather(bench.TILES_PER_GPU)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ro.step(devb, bench.TILES_PER_GPU); ro.gather(bench.TILES_PER_GPU)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
