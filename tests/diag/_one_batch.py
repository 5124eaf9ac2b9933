import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200"))
import torch, bench, gq_b200
from gq_b200 import ops
dev = torch.device("cuda:0")
reg = gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=65536, group=16, backend="cuda").to(dev).eval()
cb = reg._codebook()
z = bench.synth_latents(64, int(sys.argv[1])).to(dev)
for _ in range(3):
    ops.quantize_latent(cb, z, ops.GQ_LAYOUT_BCHW)
torch.cuda.synchronize()
