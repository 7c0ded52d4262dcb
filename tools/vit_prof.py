"""ViT-S/8 descriptor front end: per-kernel time shares (CUDA events per call) for 1024 crops.  usage: [ncu ...] python tools/vit_prof.py [B]"""
import sys, json
import torch
sys.path.insert(0, ".")
import bench
print(json.dumps(bench.descriptor_section(torch.device("cuda")), indent=1))
