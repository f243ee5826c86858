"""Layer-by-layer forward comparison of the generator program with the bf16-emulating oracle (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pcgan_oracle as O
from pcgan_b200 import networks as NW
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"
N, S = 2, 32
rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))
sd = O.make_state_dict(O.generator_keys(), 41, device=DEV)
net = NW.define_G(3, 3, 1, 64, "resnet_9blocks", "instance", init_type="normal", gpu_ids=[0])
mod = net.module
mod.load_state_dict({k: v.clone() for k, v in sd.items()})
a, _, _ = O.synthetic_batch(N, S, 300, device=DEV)
z = torch.linspace(-1, 1, N, device=DEV).view(N, 1, 1, 1)
prog = mod._program(N, S)
with torch.no_grad():
    out, ws = prog.forward(a.contiguous(), z.view(-1).contiguous())
    for tag, q in (("exact", O.Quant(False)), ("emul", O.Quant(True))):
        taps = {}
        sdq = O.make_state_dict(O.generator_keys(), 41, device=DEV)
        ref = O.generator_forward(sdq, a, z, taps=taps, q=q)
        def nhwc(buf, g):
            t = buf[: g.numel].view(g.n, g.hp, g.wp, g.c).float()
            if g.pad: t = t[:, g.pad:g.pad + g.h, g.pad:g.pad + g.w]
            return t.permute(0, 3, 1, 2)
        print("==", tag)
        print("model.1", rel(nhwc(ws.r1, prog.g_r1), taps["model.1"]))
        print("model.4", rel(nhwc(ws.r2, prog.g_r2), taps["model.4"]))
        print("model.7", rel(nhwc(ws.r3, prog.g_r3), taps["model.7"]))
        import torch.nn.functional as F
        def padded(buf, g):
            return buf[: g.numel].view(g.n, g.hp, g.wp, g.c).float().permute(0, 3, 1, 2)
        print("act.model.8 interior", rel(nhwc(ws.b[0], prog.g_b), taps["act.model.8"]), "padded", rel(padded(ws.b[0], prog.g_b), F.pad(taps["act.model.8"], (1,)*4, mode="reflect")))
        d = (nhwc(ws.b[0], prog.g_b) - taps["act.model.8"]).abs()
        print("  max abs diff", float(d.max()), "frac nonzero", float((d > 0).float().mean()), "ref absmax", float(taps["act.model.8"].abs().max()))
        print("  stats mean mine/ref", float(ws.n3.mean[0, 0]), float(taps["model.7"][0, 0].mean()), "rstd", float(ws.n3.rstd[0, 0]), float(1 / (taps["model.7"][0, 0].var(unbiased=False) + 1e-5).sqrt()))
        print("act.block10.2 interior", rel(nhwc(ws.h[0], prog.g_b), taps["act.model.10.conv_block.2"]))
        for i in range(9):
            p = "model.%d.conv_block" % (10 + i)
            print(p, rel(nhwc(ws.ra[i], prog.g_r3), taps[p + ".1"]), rel(nhwc(ws.rb[i], prog.g_r3), taps[p + ".5"]),
                  "block out", rel(nhwc(ws.b[i + 1], prog.g_b), taps["model.%d" % (10 + i)]))
        print("model.19", rel(nhwc(ws.u1r, prog.g_u1r), taps["model.19"]))
        print("model.22", rel(nhwc(ws.u2r, prog.g_u2r), taps["model.22"]))
        print("out", rel(out, ref))
