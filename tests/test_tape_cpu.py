"""Host logic of the guidance-gradient tape (theatergen_b200/backward.py::_Tape) on CPU: reverse traversal, accumulation
where a tensor has several consumers (residuals / skips), pass-through gradients that alias, nodes that must run without
an upstream gradient (the captured attention map of the last transformer).  The only device operator the tape itself
calls is `ops.add_inplace`; it is replaced by a torch add here, and the result is checked against torch autograd."""
import pytest
import torch

from theatergen_b200 import backward


@pytest.fixture()
def cpu_add(monkeypatch):
    def add_inplace(y, x):
        y.add_(x)
        return y
    monkeypatch.setattr(backward.ops, "add_inplace", add_inplace)


def test_tape_matches_autograd_on_a_residual_graph(cpu_add):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 8, generator=g)
    w1, w2 = torch.randn(8, 8, generator=g), torch.randn(8, 8, generator=g)
    tape = backward._Tape()
    # h = x W1 ; y = h W2 + h (residual: h has two consumers) ; z = y + x (x has two consumers, one a pass-through)
    h = x @ w1
    tape.add(h, lambda d: [(x, d @ w1.t())])
    y = h @ w2 + h
    tape.add(y, lambda d: [(h, d @ w2.t()), (h, d)])
    z = y + x
    tape.add(z, lambda d: [(y, d), (x, d)])
    seed = torch.randn(4, 8, generator=g)
    grads = tape.backward({z.data_ptr(): seed})
    xr = x.clone().requires_grad_(True)
    hr = xr @ w1
    ((hr @ w2 + hr) + xr).backward(seed)
    assert torch.allclose(grads[x.data_ptr()], xr.grad, atol=1e-5)
    assert torch.equal(seed, seed.clone())  # the seed (aliased as a pass-through gradient) was not modified in place


def test_pass_through_gradients_are_not_modified_in_place(cpu_add):
    x = torch.ones(3)
    tape = backward._Tape()
    a = x * 2
    tape.add(a, lambda d: [(x, 2 * d)])
    b = a + 0          # identity consumers of `a`: both hand the SAME tensor object back
    tape.add(b, lambda d: [(a, d)])
    c = a + 0
    tape.add(c, lambda d: [(a, d)])
    shared = torch.full((3,), 5.0)
    grads = tape.backward({b.data_ptr(): shared, c.data_ptr(): shared})
    assert torch.equal(shared, torch.full((3,), 5.0))
    assert torch.equal(grads[x.data_ptr()], torch.full((3,), 20.0))


def test_nodes_without_upstream_gradient(cpu_add):
    x = torch.ones(2)
    tape = backward._Tape()
    a = x + 1
    tape.add(a, lambda d: [(x, d)])
    dead = a * 3       # never reaches the loss: skipped
    calls = []
    tape.add(dead, lambda d: calls.append("dead") or [(a, 3 * d)])
    probe = a * 5      # no consumer either, but carries an injected gradient (always=True), like the captured map
    tape.add(probe, lambda d: [(a, torch.full((2,), 7.0) if d is None else d)], always=True)
    grads = tape.backward()
    assert calls == [] and torch.equal(grads[x.data_ptr()], torch.full((2,), 7.0))
