"""FusedAdam (pcgan_b200/optim.py, pcgan_adam_batched) against torch.optim.Adam, which the reference steps at
models/wsgan_emb_model.py:153-154, 451-461: same trajectory, same state names, interchangeable state_dict, LR schedulers."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _params(seed, shapes):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]


SHAPES = [(64, 4, 7, 7), (64,), (3,), (1,), (256, 256, 3, 3), (5, 3), (128, 64, 3, 3), (7,)]   # odd sizes: unaligned views


def _grads(ps, step):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(*p.shape, generator=g).to(DEV) * (0.1 + 0.01 * i) for i, p in enumerate(ps)]


def test_matches_torch_adam_over_steps_and_lr_schedule():
    from pcgan_b200.optim import FusedAdam
    pa, pb = _params(0, SHAPES), _params(0, SHAPES)
    oa = torch.optim.Adam(pa, lr=2e-4, betas=(0.5, 0.999))
    ob = FusedAdam(pb, lr=2e-4, betas=(0.5, 0.999))
    sa = torch.optim.lr_scheduler.LambdaLR(oa, lr_lambda=lambda e: 1.0 - 0.2 * e)
    sb = torch.optim.lr_scheduler.LambdaLR(ob, lr_lambda=lambda e: 1.0 - 0.2 * e)
    for step in range(12):
        for ps in (pa, pb):
            for p, g in zip(ps, _grads(ps, step)):
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
        oa.step()
        ob.step()
        if step % 4 == 3:
            sa.step()
            sb.step()
            assert ob.param_groups[0]["lr"] == pytest.approx(oa.param_groups[0]["lr"])
    for a, b in zip(pa, pb):
        assert float((a - b).abs().max()) <= 2e-6 * (1.0 + float(a.abs().max()))     # fp32 rounding of a different op order
    assert all(b._version >= 12 for b in pb)       # in-place updates are visible to autograd / version-keyed caches
    st = ob.state[pb[0]]
    assert set(st) >= {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 12.0
    assert torch.allclose(st["exp_avg"], oa.state[pa[0]]["exp_avg"], rtol=1e-5, atol=1e-8)
    assert torch.allclose(st["exp_avg_sq"], oa.state[pa[0]]["exp_avg_sq"], rtol=2e-5, atol=1e-10)


def test_state_dict_interchanges_with_torch_adam():
    from pcgan_b200.optim import FusedAdam
    pa, pb, pc = _params(1, SHAPES), _params(1, SHAPES), _params(1, SHAPES)
    oa = torch.optim.Adam(pa, lr=1e-3, betas=(0.5, 0.999))
    ob = FusedAdam(pb, lr=1e-3, betas=(0.5, 0.999))
    for step in range(3):
        for ps, o in ((pa, oa), (pb, ob)):
            for p, g in zip(ps, _grads(ps, step)):
                p.grad = g.clone()
            o.step()
    # fused -> torch
    oc = torch.optim.Adam(pc, lr=1e-3, betas=(0.5, 0.999))
    oc.load_state_dict(ob.state_dict())
    for p, q in zip(pc, pb):
        p.data.copy_(q.data)
    # torch -> fused
    od = FusedAdam(pa, lr=1e-3, betas=(0.5, 0.999))
    od.load_state_dict(oa.state_dict())
    for step in range(3, 6):
        for ps, o in ((pc, oc), (pa, od), (pb, ob)):
            for p, g in zip(ps, _grads(ps, step)):
                p.grad = g.clone()
            o.step()
    for a, b, c in zip(pa, pb, pc):
        tol = 2e-6 * (1.0 + float(b.abs().max()))
        assert float((a - b).abs().max()) <= tol and float((c - b).abs().max()) <= tol


def test_refuses_what_it_does_not_implement():
    from pcgan_b200.optim import FusedAdam
    ps = _params(2, [(4, 4)])
    with pytest.raises(NotImplementedError):
        FusedAdam(ps, weight_decay=0.1)
    with pytest.raises(NotImplementedError):
        FusedAdam(ps, amsgrad=True)
    o = FusedAdam(ps)
    with pytest.raises(RuntimeError, match="no gradient"):
        o.step()
    cpu = [torch.nn.Parameter(torch.zeros(3))]
    cpu[0].grad = torch.zeros(3)
    with pytest.raises(RuntimeError, match="CUDA"):
        FusedAdam(cpu).step()


def test_capturable_in_a_cuda_graph():
    from pcgan_b200.optim import FusedAdam
    pa, pb = _params(3, SHAPES), _params(3, SHAPES)
    lr = torch.tensor(2e-4, device=DEV)
    oa = torch.optim.Adam(pa, lr=2e-4, betas=(0.5, 0.999))
    ob = FusedAdam(pb, lr=lr, betas=(0.5, 0.999))
    gbuf = [torch.zeros_like(p) for p in pb]
    for p, g in zip(pb, gbuf):
        p.grad = g
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for g, v in zip(gbuf, _grads(pb, 0)):
            g.copy_(v)
        ob.step()                                 # eager warm-up builds the pointer table
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):    # capture only records: the first replayed update is step 1
            ob.step()
    torch.cuda.current_stream().wait_stream(side)
    for step in range(1, 5):
        for g, v in zip(gbuf, _grads(pb, step)):
            g.copy_(v)
        if step == 3:
            lr.fill_(1e-4)
        graph.replay()
    for step in range(5):
        for p, g in zip(pa, _grads(pa, step)):
            p.grad = g.clone()
        oa.param_groups[0]["lr"] = 1e-4 if step >= 3 else 2e-4
        oa.step()
    torch.cuda.synchronize()
    for a, b in zip(pa, pb):
        assert float((a - b).abs().max()) <= 2e-6 * (1.0 + float(a.abs().max()))
    assert float(ob.state[pb[0]]["step"]) == 5.0
