"""socioreasoner_b200 -- B200-native (sm_100a) implementation of SocioReasoner's data-parallel VLM hot path.

Package layout (only what the path needs):
  csrc/          hand-written CUDA kernels + the C ABI (include/socioreasoner_b200.h)
  lib.py         ctypes binding (raw device pointers + stream; no fallback)
  plan.py        host-side index planning (ViT windows, attention work lists, M-RoPE ids)
  weights.py     HF-checkpoint-layout -> kernel-layout weight packing
  engine.py      Qwen2.5-VL forward (ViT -> merger -> scatter -> decoder -> logits), greedy decode, mask epilogue
  modeling.py    HF-style module surface (forward(...).logits / generate) the reference's strategies call
  roll_shim/     mirror of the reference's roll.* strategy / provider / DataProto surface for this path
  rollout.py     data-parallel rollout over tiles + NCCL all-gather of masks / metrics
"""
__version__ = "0.1.0"
