import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "diffusion-models-made-easy_b200"))
from dmme_b200 import ops
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for c, heads in ((256, 4), (128, 4)):
    n, L = 256, 256
    dh = c // heads
    qkv = torch.randn(n, L, 3 * c, device=dev, generator=g).bfloat16()
    flat = qkv.view(-1)
    out = torch.empty(n, L, c, device=dev, dtype=torch.bfloat16)
    for rep in range(3):
        ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, L, dh, c ** -0.5, L * 3 * c, 3 * c, 3 * dh, False, 0, True, out)
    torch.cuda.synchronize()
