"""the final-loss pass and get_adv_data share ONE field build per
sign and one theta build (adv_compose_solver.py:236-279, 435-463 rebuild them per call), get_adv_data's
outputs are value-checked against the oracle with the parameters the call ended on, and the AdvBias
options no other GPU test touches (space='linear', init_mode gaussian / identity)."""
import pytest
import torch

from oracle import advchain_oracle as orc
from tests.golden.cases import CASES, stage_cfgs
from tests.helpers import cuda_solver, oracle_solver, rel_err

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _model(case, dev):
    torch.manual_seed(3)
    conv = torch.nn.Conv2d if case["d"] == 2 else torch.nn.Conv3d
    return conv(case["size"][1], case["K"], 3, 1, 1).eval().to(dev)


def _smooth_data(size):
    """A smooth volume (low-pass noise): resampling it is well conditioned, so end-to-end values can be held
    to the strict output tolerance (white noise cannot: tests/test_gpu_golden.py)."""
    d = len(size) - 2
    torch.manual_seed(17)
    small = torch.rand(size[0], size[1], *[max(3, s // 8) for s in size[2:]])
    mode = "bilinear" if d == 2 else "trilinear"
    return torch.nn.functional.interpolate(small, size=tuple(size[2:]), mode=mode, align_corners=True)


def _oracle_with_params(case, sol):
    osol = oracle_solver(case)
    for s, t in zip(osol.stages, sol.chain_of_transforms):
        s.init()
        s.param = t.param.detach().cpu().clone()
        s.eval()
    return osol


@pytest.mark.parametrize("name", ["c2d_full", "c3d_full"])
def test_final_pass_builds_each_field_once(name):
    """calc_adv_consistency_loss runs forward, predict_forward (mask), predict_backward twice (mask and
    prediction): the reference builds phi(+v) twice and phi(-v) twice (adv_compose_solver.py:253-271); here
    every squaring step runs once per sign, and a second pass with unchanged parameters builds nothing."""
    from advchain_b200 import _lib
    dev = _dev()
    case = CASES[name]
    sol = cuda_solver(case, dev)
    model = _model(case, dev)
    data = _smooth_data(case["size"]).to(dev)
    init = sol.get_init_output(model, data)
    sol.init_random_transformation()
    morph = [t for t in sol.chain_of_transforms if t.get_name() == "morph"][0]
    nb = morph._nb_steps()
    _lib.launch_count(reset=True)
    loss1, adv1, out1, warped1 = sol.calc_adv_consistency_loss(data, model, init)
    torch.cuda.synchronize()
    assert _lib.launch_count("ss_step") == 2 * nb, (_lib.launch_count("ss_step"), nb)
    assert _lib.launch_count("affine_theta_fwd") == 1
    # same parameters again: every field / theta comes from the cache, values identical
    _lib.launch_count(reset=True)
    loss2, adv2, out2, warped2 = sol.calc_adv_consistency_loss(data, model, init)
    torch.cuda.synchronize()
    assert _lib.launch_count("ss_step") == 0 and _lib.launch_count("affine_theta_fwd") == 0
    assert torch.equal(adv1, adv2) and torch.equal(warped1, warped2) and float(loss1) == float(loss2)
    # a fresh solver that cannot have anything cached gives the same values (cached == uncached path)
    sol2 = cuda_solver(case, dev)
    for t, u in zip(sol.chain_of_transforms, sol2.chain_of_transforms):
        u.init_parameters()
        u.param = t.param.detach().clone()
    loss3, adv3, out3, warped3 = sol2.calc_adv_consistency_loss(data, model, init)
    assert torch.equal(adv1, adv3) and torch.equal(warped1, warped3)
    assert abs(float(loss1) - float(loss3)) <= 1e-6 * abs(float(loss3))
    # a parameter update invalidates the cache
    with torch.no_grad():
        morph.param = morph.param * 0.5
    _lib.launch_count(reset=True)
    sol.calc_adv_consistency_loss(data, model, init)
    torch.cuda.synchronize()
    assert _lib.launch_count("ss_step") == 2 * morph._nb_steps()


@pytest.mark.parametrize("name", ["c2d_full", "c3d_full"])
@pytest.mark.parametrize("n_iter", [0, 1])
def test_get_adv_data_values(name, n_iter):
    """get_adv_data (adv_compose_solver.py:435-463): augmented data = forward(data), augmented label =
    predict_forward(init_output), both with the transforms the call ended on.  Checked by value against the
    oracle evaluated with those parameters, on a smooth volume at the strict tolerance; the label pass and the
    data pass share one phi(+v) and one theta."""
    from advchain_b200 import _lib
    dev = _dev()
    case = CASES[name]
    sol = cuda_solver(case, dev)
    model = _model(case, dev)
    data = _smooth_data(case["size"])
    init = sol.get_init_output(model, data.to(dev))
    torch.manual_seed(5)
    _lib.launch_count(reset=True)
    aug, label = sol.get_adv_data(data.to(dev), model, init_output=init, n_iter=n_iter)
    torch.cuda.synchronize()
    morph = [t for t in sol.chain_of_transforms if t.get_name() == "morph"][0]
    if n_iter == 0:
        assert _lib.launch_count("ss_step") == morph._nb_steps()       # phi(+v) only, once
    assert aug.shape == data.shape and label.shape == init.shape
    assert all(not t.is_training for t in sol.chain_of_transforms) or n_iter == 0
    osol = _oracle_with_params(case, sol)
    with torch.no_grad():
        ref_aug = osol.forward(data)
        ref_label = osol.predict_forward(init.detach().cpu())
    assert rel_err(aug, ref_aug) < OUT_TOL, rel_err(aug, ref_aug)
    assert rel_err(label, ref_label) < 5 * OUT_TOL, rel_err(label, ref_label)   # logits of a random conv: rougher than the image


@pytest.mark.parametrize("d,size", [(2, [2, 1, 48, 64]), (3, [1, 2, 16, 24, 32])])
@pytest.mark.parametrize("space", ["linear", "log"])
@pytest.mark.parametrize("init_mode", ["random", "gaussian", "identity"])
def test_bias_space_and_init_modes(d, size, space, init_mode):
    """AdvBias options (adv_bias.py:84-137, 176-188): multiplicative field in 'linear' space (1 + upsampled
    lattice instead of exp), init_mode gaussian / identity (unbounded control points: rescale does not clip)."""
    from advchain_b200.augmentor import AdvBias
    dev = _dev()
    cfg = dict(stage_cfgs(d, size)["bias"], space=space, init_mode=init_mode)
    torch.manual_seed(9)
    t = AdvBias(d, cfg, device=dev)
    t.init_parameters()
    o = orc.Bias(cfg)
    o.init()
    assert tuple(o.param.shape) == tuple(t.param.shape)
    if init_mode == "identity":
        assert float(t.param.abs().max()) == 0.0
    cp = (torch.rand_like(o.param) - 0.5) * 1.2          # part of the field leaves the clip range
    x = torch.rand(*size)
    gout = torch.randn(*size)
    o.param = cp.clone().requires_grad_(True)
    x0 = x.clone().requires_grad_(True)
    ref = o.fwd(x0)
    ref.backward(gout)
    t.param = cp.to(dev).requires_grad_(True)
    x1 = x.to(dev).requires_grad_(True)
    out = t.forward(x1)
    out.backward(gout.to(dev))
    assert rel_err(out, ref) < OUT_TOL
    assert rel_err(x1.grad, x0.grad) < OUT_TOL
    assert rel_err(t.param.grad, o.param.grad) < 1e-4
    # rescale_parameters: clip to the init range for 'random', unbounded otherwise (adv_bias.py:150)
    o.param = (cp * 3).clone()
    o.rescale()
    t.param = (cp * 3).to(dev)
    t.rescale_parameters()
    assert rel_err(t.param, o.param) < 1e-6
