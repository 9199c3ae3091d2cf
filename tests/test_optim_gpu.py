"""GPU parity: FlatAdam (mm_clip_adam_step: clip_grad_norm_ + Adam.step on flat buffers) vs torch.nn.utils.clip_grad_norm_ +
torch.optim.Adam on the CPU in float32 -- the pair the reference runs after every backward pass
(offpolicy/algorithms/qmix/qmix.py:200-208).  Same operation order, so parameters agree to float32 rounding: 2e-6 relative
after 25 steps; the returned norm to 1e-6 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 47), (64,), (64, 64), (192, 64), (192,), (5, 64), (5,), (1,), (3, 7, 11)]


def _models(cuda, seed):
    g = torch.Generator().manual_seed(seed)
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES]
    mine = [torch.nn.Parameter(p.detach().clone().to(cuda)) for p in ref]
    return ref, mine, g


@pytest.mark.parametrize("max_norm,weight_decay", [(10.0, 0.0), (0.5, 0.0), (0.0, 0.0), (0.5, 1e-2)])
def test_flat_adam_matches_torch_adam(cuda, max_norm, weight_decay):
    from minimarl_b200.optim import FlatAdam
    ref, mine, g = _models(cuda, 3)
    o_ref = torch.optim.Adam(ref, lr=5e-4, eps=1e-5, weight_decay=weight_decay)
    o_mine = FlatAdam(mine, lr=5e-4, eps=1e-5, weight_decay=weight_decay)
    for it in range(25):
        o_ref.zero_grad()
        o_mine.zero_grad()
        scale = 10.0 if it % 3 == 0 else 0.01                        # steps that clip and steps that do not
        for pr, pm in zip(ref, mine):
            gr = torch.randn(pr.shape, generator=g) * scale
            pr.grad = gr.clone()
            pm.grad.add_(gr.to(cuda))                                 # autograd accumulates in place into the bucket's views
        if max_norm:
            n_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        else:
            n_ref = torch.linalg.vector_norm(torch.cat([p.grad.reshape(-1) for p in ref]))
        o_ref.step()
        n_mine = o_mine.clip_and_step(max_norm)
        assert abs(float(n_mine) - float(n_ref)) < 1e-6 * float(n_ref)
    for pr, pm in zip(ref, mine):
        assert (pm.detach().cpu() - pr.detach()).abs().max().item() < 2e-6 * max(1.0, pr.detach().abs().max().item())
    sd_ref, sd = o_ref.state_dict(), o_mine.state_dict()
    for k in sd_ref["state"]:
        assert float(sd["state"][k]["step"]) == float(sd_ref["state"][k]["step"]) == 25
        assert torch.allclose(sd["state"][k]["exp_avg"].cpu(), sd_ref["state"][k]["exp_avg"], rtol=1e-4, atol=1e-5)
        assert torch.allclose(sd["state"][k]["exp_avg_sq"].cpu(), sd_ref["state"][k]["exp_avg_sq"], rtol=1e-4, atol=1e-6)


def test_flat_adam_views_state_and_replaced_grads(cuda):
    """Parameters are views of one buffer on 512-byte boundaries; replaced ``p.grad`` tensors are packed back; a state dict
    round-trips into a fresh optimiser and both continue identically; autograd writes through the views."""
    from minimarl_b200.optim import FlatAdam
    _, mine, g = _models(cuda, 5)
    opt = FlatAdam(mine, lr=1e-3, eps=1e-8)
    assert all(p.data_ptr() % 512 == 0 and p.grad.data_ptr() % 512 == 0 for p in mine)
    loss = sum((p ** 2).sum() for p in mine)
    opt.zero_grad()
    loss.backward()
    assert opt.bucket.attached()
    for p in mine:
        assert torch.equal(p.grad, 2 * p.detach())
    before = [p.detach().clone() for p in mine]
    versions = [p._version for p in mine]
    mine[0].grad = torch.ones_like(mine[0])                          # replaced: not a view any more
    opt.clip_and_step(1.0)
    assert all(p._version > v for p, v in zip(mine, versions))       # caches keyed on p._version see the raw-pointer update
    assert opt.bucket.attached() and all(not torch.equal(b, p.detach()) for b, p in zip(before, mine))
    sd = opt.state_dict()
    clone = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    opt2 = FlatAdam(clone, lr=1e-3, eps=1e-8)
    opt2.load_state_dict(sd)
    for o, ps in ((opt, mine), (opt2, clone)):
        o.zero_grad()
        for i, p in enumerate(ps):
            p.grad.add_(float(i + 1))
        o.clip_and_step(0.0)
    for a, b in zip(mine, clone):
        assert torch.equal(a.detach(), b.detach())


def test_flat_adam_replays_in_a_cuda_graph(cuda):
    """The step counter lives on the device: a captured step replays with advancing bias correction."""
    from minimarl_b200.optim import FlatAdam
    ref, mine, g = _models(cuda, 9)
    o_ref = torch.optim.Adam(ref, lr=1e-3, eps=1e-8)
    o_mine = FlatAdam(mine, lr=1e-3, eps=1e-8)
    grads = [torch.randn(p.shape, generator=g) for p in ref]
    for pm, gr in zip(mine, grads):
        pm.grad.copy_(gr.to(cuda))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        o_mine.clip_and_step(1.0)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o_mine.clip_and_step(1.0)
    for _ in range(3):
        graph.replay()
    for _ in range(4):                                               # the warm-up step + three replays (a capture runs nothing)
        for pr, gr in zip(ref, grads):
            pr.grad = gr.clone()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step()
    for pr, pm in zip(ref, mine):
        assert (pm.detach().cpu() - pr.detach()).abs().max().item() < 2e-6 * max(1.0, pr.detach().abs().max().item())
