#!/usr/bin/env python
"""Small end-to-end pass over every kernel of the library (fast / robust search at g = 4, 8, 16, the
latent path, dequant from the packed stream, usage histogram, KL penalty), meant to be run under
compute-sanitizer:

    compute-sanitizer --tool memcheck  python tests/diag/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tests/diag/sanitize_smoke.py
    compute-sanitizer --tool synccheck python tests/diag/sanitize_smoke.py

Round 1 on B200: 0 errors / 0 hazards with all three tools; indices equal to the oracle's."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "vq-vae-from-gaussian-vae_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, gq_b200, oracle
from gq_b200_host import host_codebook
dev=torch.device("cuda:0")
for (N,g,M) in ((1000,16,300),(4096,4,700),(4096,8,300)):
    cbn,nlp=host_codebook(N,g)
    cb=gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    gen=torch.Generator().manual_seed(1)
    mu=torch.randn(M,g,generator=gen)*0.8; std=torch.exp(0.5*(-1.2+0.4*torch.randn(M,g,generator=gen)))
    idx,zh=gq_b200.quantize_rows(cb, mu.to(dev), std.to(dev), std.log().to(dev))
    o,_,_,_=oracle.quantize(mu.numpy(),std.numpy(),std.log().numpy(),cbn,nlp)
    print(N,g,M,"mismatch",int((idx.cpu().numpy()!=o).sum()))
    cb.set_mode("robust")
    idx,zh=gq_b200.quantize_rows(cb, mu.to(dev), std.to(dev), std.log().to(dev))
    print(" robust mismatch",int((idx.cpu().numpy()!=o).sum()))
# sharp rows (pruned exact scan inside settle_kernel), one NaN row (sliced full scan), and a many-segment plan
# (small M against a larger codebook: rescore_many_kernel settles and prunes inline)
for (N,g,M) in ((4096,16,200),(16384,16,64)):
    cbn,nlp=host_codebook(N,g)
    cb=gq_b200.Codebook(torch.from_numpy(cbn).to(dev), torch.from_numpy(nlp).to(dev), 1.0)
    gen=torch.Generator().manual_seed(2)
    mu=torch.randn(M,g,generator=gen); std=torch.abs(torch.randn(M,g,generator=gen)).clamp_min(1e-4)
    std[::7,3]=2e-4
    o,_,_,_=oracle.quantize(mu.numpy(),std.numpy(),std.log().numpy(),cbn,nlp)
    for mode in ("fast","robust"):
        cb.set_mode(mode)
        st=torch.zeros(8,dtype=torch.int32,device=dev)
        idx,zh=gq_b200.quantize_rows(cb, mu.to(dev), std.to(dev), std.log().to(dev), stats=st)
        print(N,g,M,mode,"stress mismatch",int((idx.cpu().numpy()!=o).sum()),"uncertified rows",int(st[2]))
    mu2=mu.clone(); mu2[5,2]=float("nan")
    idx,zh=gq_b200.quantize_rows(cb, mu2.to(dev), std.to(dev), std.log().to(dev))
    print(" nan row index", int(idx[5]))
# latent path + KL + tokens
reg=gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=4096, group=16, backend="cuda").cuda().eval()
z=torch.randn(2,32,8,8,device=dev)*0.5
zhat,info=reg(z)
p=reg.pack_indices(info["indices"]); h=reg.accumulate_usage(info["indices"]); print(reg.codebook_usage(h)[0].item(), torch.equal(reg.dequant(p), zhat))
# latent call whose rows need the final-layout pass, with rows on the exact-scan list (clamped logvar extremes), g = 4 and 16
for grp in (16, 4):
    rg=gq_b200.GaussianQuantRegularizer(format="bchw", n_samples=4096, group=grp, backend="cuda").cuda().eval()
    zz=torch.randn(3,32,8,8,device=dev)*0.6; zz[0,16:,0,0]=25.0; zz[0,16:,0,1]=-40.0; zz[1,16:,3,3]=-40.0
    zq,inf=rg(zz)
    print(" tail-merge latent call, group", grp, "dequant == zhat", torch.equal(rg.dequant(inf["indices"]), zq))
reg.train(); zhat,info=reg(z); print(float(info["kl_loss"]))
torch.cuda.synchronize(); print("done")
