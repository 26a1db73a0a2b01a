"""GPU parity tests of the fused eval tower (csrc/tower_eval.cu, dcnr_tower_eval; pytest -m gpu).

Checker: the float64 oracle's deep tower (oracle/dcnr_oracle.forward(..., return_parts=True): initial layer, ResBlocks
with running-statistics BatchNorm, deep half of the final dot -- main.py:83-90,120-127 in eval()).  Contract: fp16x3
logits within 1e-5 (max-abs-normalised) like every parity mode; bf16 is a stated-tolerance mode and the test states it.
Covered: ragged row counts (1, 127, 128, 129, ...), many tiles per CTA (tile-to-tile
pipelining), 1 / 2 / 4 ResBlocks, a 96-wide padded input (three K chunks in the initial layer), the whole-model eval
call, the out-of-range-id flag and the fp16-range fallback.
"""
import numpy as np
import pytest
import torch

from oracle import dcnr_oracle as orc
from tests.helpers import synth_inputs

pytestmark = pytest.mark.gpu

TOL = 1e-5
BF16_TOL = 3e-2          # stated tolerance of the bf16 mode (measured ~5e-3 on these states; SURVEY 8d predicted 1.4e-3..1e-2)
CAT = {"city": 100, "hotel_type": 6}


def _state(R=2, emb=16, n_users=3000, n_items=1200, seed=11):
    params = dict(emb_dim=emb, hidden_dim=256, n_cross_layers=3, n_res_blocks=R, dropout=0.0)
    st = orc.make_state(n_users, n_items, CAT, 11, params, seed=seed, emb_scale=0.1, randomize_bn=True)
    return params, st, n_users, n_items


def _model(params, st, n_users, n_items, precision):
    import dcnr_b200
    m = dcnr_b200.DCN_RecSys(n_users, n_items, CAT, 11, params, precision=precision)
    m.load_state_dict(st)
    return m.cuda().eval()


def _tower_call(m, x0p, cross, precision, options):
    from dcnr_b200 import _cabi as C
    dims, ps = m._dims(), m._param_struct()
    B = x0p.shape[0]
    out = torch.full((B,), float("nan"), device="cuda")
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    ws = torch.empty(C.lib().dcnr_tower_eval_workspace_bytes(dims), dtype=torch.uint8, device="cuda")
    C.check(C.lib().dcnr_tower_eval(dims, ps, C.ptr(x0p), x0p.shape[1], C.ptr(cross), C.ptr(out), B, C.PRECISIONS[precision],
                                    options, C.ptr(flags), C.ptr(ws), ws.numel(), C.stream()))
    torch.cuda.synchronize()
    return out.cpu(), int(flags[0].item())


def _oracle_deep(st, u, i, c, x):
    st64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in st.items()}
    logits, parts = orc.forward(st64, u, i, c, x.double(), training=False, return_parts=True)
    H = parts["deep"].shape[1]
    deep = parts["deep"] @ st64["final_linear.weight"][0, :H] + st64["final_linear.bias"][0]
    return logits.reshape(-1), deep, parts["x0"]


@pytest.mark.parametrize("B", [1, 127, 128, 129, 300, 4096, 37_001, 100_000])
def test_tower_operator_matches_oracle(B, options=0):
    params, st, nu, ni = _state()
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, B, seed=5 + B)
    _, deep, x0 = _oracle_deep(st, u, i, c, x)
    m = _model(params, st, nu, ni, "fp16x3")
    Dp = m._dims().in_dim_pad
    x0p = torch.zeros(B, Dp)
    x0p[:, : x0.shape[1]] = x0.float()
    cross = torch.randn(B, generator=torch.Generator().manual_seed(1))
    got, flags = _tower_call(m, x0p.cuda(), cross.cuda(), "fp16x3", options)
    assert flags == 0
    assert orc.max_abs_normalised(got, deep + cross.double()) < TOL


@pytest.mark.parametrize("options", [1 << 8, 2 << 8, 3 << 8, 7 << 8], ids=["1_cta", "2_ctas", "3_ctas", "7_ctas"])
def test_tower_many_tiles_per_cta(options):
    """Grid capped to 1..6 CTAs: every CTA walks many tiles (the tile-to-tile hand-over of the TMEM buffers and the ring)."""
    params, st, nu, ni = _state()
    B = 128 * 23 + 17
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, B, seed=31)
    _, deep, x0 = _oracle_deep(st, u, i, c, x)
    m = _model(params, st, nu, ni, "fp16x3")
    x0p = torch.zeros(B, m._dims().in_dim_pad)
    x0p[:, : x0.shape[1]] = x0.float()
    got, flags = _tower_call(m, x0p.cuda(), None, "fp16x3", options)
    assert flags == 0
    assert orc.max_abs_normalised(got, deep) < TOL


@pytest.mark.parametrize("R,emb", [(1, 16), (4, 16), (2, 32), (3, 48)])
def test_tower_depths_and_input_widths(R, emb):
    """1..4 ResBlocks; emb 32 -> D = 89 -> three 32-wide K chunks in the initial layer, emb 48 -> D = 121 -> four."""
    params, st, nu, ni = _state(R=R, emb=emb, seed=3 + R)
    B = 5000
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, B, seed=77)
    _, deep, x0 = _oracle_deep(st, u, i, c, x)
    m = _model(params, st, nu, ni, "fp16x3")
    Dp = m._dims().in_dim_pad
    x0p = torch.zeros(B, Dp)
    x0p[:, : x0.shape[1]] = x0.float()
    for options in (0, 5 << 8):
        got, flags = _tower_call(m, x0p.cuda(), None, "fp16x3", options)
        assert flags == 0
        assert orc.max_abs_normalised(got, deep) < TOL, (R, emb, options)


@pytest.mark.parametrize("precision", ["fp16x3", "bf16"])
def test_whole_model_eval_uses_fused_tower(precision):
    """DCN_RecSys.eval() forward (gather + cross + fused tower) against the float64 oracle and against the tf32x3 path."""
    from dcnr_b200 import _cabi as C
    params, st, nu, ni = _state()
    B = 20_000
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, B, seed=9, zipf=True)
    ref, _, _ = _oracle_deep(st, u, i, c, x)
    m = _model(params, st, nu, ni, precision)
    C.launch_count(reset=True)
    with torch.no_grad():
        out = m(u.cuda(), i.cuda(), c.cuda(), x.cuda())
    launches = C.launch_count()
    assert launches == 3, launches          # weight pack, gather + cross, fused tower
    C.launch_count(reset=True)
    with torch.no_grad():
        again = m(u.cuda(), i.cuda(), c.cuda(), x.cuda())
    assert C.launch_count() == 2            # the prepared weights are cached until a parameter changes
    assert torch.equal(again, out)
    err = orc.max_abs_normalised(out.cpu(), ref)
    assert err < (TOL if precision == "fp16x3" else BF16_TOL), err
    if precision == "bf16":
        assert err > 1e-5                   # it really is the reduced-precision arithmetic (not an alias of a parity mode)


def test_prepared_weights_follow_parameter_updates():
    """The cached weight pack of the fused tower is rebuilt after in-place updates by torch (an optimizer step), by this
    library's fused Adam, and after a training step of our kernels moved the BatchNorm running statistics."""
    import dcnr_b200
    from dcnr_b200 import functional as F_
    params, st, nu, ni = _state()
    B = 3000
    u, i, c, x, y = synth_inputs(nu, ni, CAT, 11, B, seed=4)
    args = (u.cuda(), i.cuda(), c.cuda(), x.cuda())
    m = _model(params, st, nu, ni, "fp16x3")

    def fresh_eval():
        f = dcnr_b200.DCN_RecSys(nu, ni, CAT, 11, params, precision="fp16x3")
        f.load_state_dict(m.state_dict())
        f = f.cuda().eval()
        with torch.no_grad():
            return f(*args)

    with torch.no_grad():
        first = m(*args)
        m.res_blocks[0].layer1.weight.mul_(1.5)                          # torch in-place update
        assert torch.equal(m(*args), fresh_eval()) and not torch.equal(m(*args), first)
        w = m.res_blocks[1].layer2.weight
        F_.adam_step_(w, torch.ones_like(w), torch.zeros_like(w), torch.zeros_like(w), 1, 0.05)      # raw-pointer update
        assert torch.equal(m(*args), fresh_eval())
    m.train()
    m(*args).sum().backward()                                            # moves the running statistics
    m.eval()
    with torch.no_grad():
        assert torch.equal(m(*args), fresh_eval())


def test_eval_reports_out_of_range_ids():
    params, st, nu, ni = _state()
    m = _model(params, st, nu, ni, "fp16x3")
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, 64, seed=1)
    u[7] = nu + 5
    with pytest.raises(IndexError):
        with torch.no_grad():
            m(u.cuda(), i.cuda(), c.cuda(), x.cuda())
    u[7] = 0
    with torch.no_grad():
        m(u.cuda(), i.cuda(), c.cuda(), x.cuda())       # the flag was cleared: the next call is clean
    # train-mode forwards record the same flag (no read-back per step); check_eval_flags() raises at the caller's pace
    m.train()
    i[3] = ni + 1
    m(u.cuda(), i.cuda(), c.cuda(), x.cuda())
    with pytest.raises(IndexError):
        m.check_eval_flags()
    assert m.check_eval_flags() is False


def test_fp16_range_overflow_falls_back_to_tf32x3():
    """Activations beyond the fp16 range of the fused tower (initial layer scaled by 1e5) set the range flag; the module
    re-runs the batch on the tf32x3 kernels, so the result still meets the 1e-5 contract."""
    params, st, nu, ni = _state()
    st = {k: v.clone() for k, v in st.items()}
    st["initial_deep_layer.weight"] *= 1e5
    B = 700
    u, i, c, x, _ = synth_inputs(nu, ni, CAT, 11, B, seed=2)
    ref, deep, x0 = _oracle_deep(st, u, i, c, x)
    m = _model(params, st, nu, ni, "fp16x3")
    x0p = torch.zeros(B, m._dims().in_dim_pad)
    x0p[:, : x0.shape[1]] = x0.float()
    _, flags = _tower_call(m, x0p.cuda(), None, "fp16x3", 0)
    assert flags & 2
    with torch.no_grad():
        out = m(u.cuda(), i.cuda(), c.cuda(), x.cuda())
    assert orc.max_abs_normalised(out.cpu(), ref) < TOL
    assert np.isfinite(out.cpu().numpy()).all()
