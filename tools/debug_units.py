"""Per-unit forward comparison of a generator engine against torch running the module's own layers."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import randomize, rel
from tactile_gan_b200.generators.generators import create_gen

kind, nf, size, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
net = create_gen(kind, 3, 3, nf, True)
randomize(net)
net = net.cuda()
x = (torch.rand(n, 3, size, size, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
with torch.no_grad():
    out = net(x)
eng = net._engine(n, size, size, False)
def nchw(act): return act.buf[..., :act.c].permute(0, 3, 1, 2).float()
with torch.no_grad():
    for u in eng.units:
        srcs = [nchw(t) for t in u.srcs]
        xin = torch.cat(srcs, 1)
        l = u.layer
        w = l.weight.detach().bfloat16().float()
        b = l.bias.detach() if l.bias is not None else None
        if l.kind == "conv":
            z = F.conv2d(xin, w, b, stride=l.stride, padding=l.pad)
        else:
            z = F.conv_transpose2d(xin, w, b, stride=l.stride, padding=l.pad)
        if u.norm:
            e_raw = rel(u.raw[..., :l.O].permute(0, 3, 1, 2), z)
            zz = u.raw[..., :l.O].permute(0, 3, 1, 2).float()
            g = u.gamma.detach() if u.gamma is not None else None
            bb = u.beta.detach() if u.beta is not None else None
            y = F.relu(F.instance_norm(zz, weight=g, bias=bb, eps=1e-5))
            mean = zz.mean((2, 3)); var = zz.var((2, 3), unbiased=False)
            e_mean = rel(u.mr[:, :l.O, 0], mean); e_rstd = rel(u.mr[:, :l.O, 1], (var + 1e-5).rsqrt())
        else:
            e_raw = -1; y = z; e_mean = e_rstd = -1
        e_y = rel(nchw(u.y), y)
        extra = ""
        if u.pool is not None:
            pr = F.max_pool2d(nchw(u.y), 2) if u.pool_mode == 2 else F.avg_pool2d(nchw(u.y), 2)
            extra += f" pool={rel(nchw(u.pool), pr):.2e}"
        print(f"{u.name:12s} {l.kind:5s} k{l.kh}s{l.stride} in{l.in_split} out{l.O} {tuple(u.y.buf.shape)} raw={e_raw:.2e} mean={e_mean:.2e} rstd={e_rstd:.2e} y={e_y:.2e}{extra}")
