#!/usr/bin/env python
"""Long-run sanity of the public samplers on random-init weights: full 1000-step DDPM / IDDPM chains, 50-step DDIM and the
GenerateImage history, checked for finite values.  usage: python tools/sanity_generate.py"""
import sys, time, torch
sys.path.insert(0, "diffusion-models-made-easy_b200")
import dmme_b200
from dmme_b200.models.ddpm import UNet
from dmme_b200.models import iddpm
torch.manual_seed(0)
dev="cuda"
ddpm = dmme_b200.DDPM(UNet().eval(), 1000).to(dev)
t0=time.time(); x = ddpm.generate((256,3,32,32), seed=1); torch.cuda.synchronize(); t1=time.time()
print("DDPM 1000 steps x256: %.2f s, finite %s, mean %.3f std %.3f absmax %.2f" % (t1-t0, bool(torch.isfinite(x).all()), float(x.mean()), float(x.std()), float(x.abs().max())))
ddim = dmme_b200.DDIM(ddpm.model, 1000, 50).to(dev)
x = ddim.generate((256,3,32,32), seed=1); torch.cuda.synchronize()
print("DDIM 50 steps: finite %s std %.3f" % (bool(torch.isfinite(x).all()), float(x.std())))
idd = dmme_b200.IDDPM(iddpm.UNet().eval(), 1000).to(dev)
t0=time.time(); x = idd.generate((64,3,32,32), seed=1); torch.cuda.synchronize(); t1=time.time()
print("IDDPM 1000 steps x64: %.2f s finite %s std %.3f" % (t1-t0, bool(torch.isfinite(x).all()), float(x.std())))
h = ddpm.generate_history((8,3,32,32), vis_length=20, seed=3)
print("history", tuple(h.shape), float(h.min()), float(h.max()))
