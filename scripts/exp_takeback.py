"""GPU-box diagnostic of the speculative loop's take-back path: the scenario of
tests/test_gpu_golden.py::test_graph_loop_takes_back_a_mispredicted_iteration, several seeds, with the step counts
the eager loop used, the predictions of the graph loop and the difference of the results."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advchain_b200.augmentor import AdvMorph, ComposeAdversarialTransformSolver, _ops  # noqa: E402
from tests.golden.cases import stage_cfgs  # noqa: E402

dev = torch.device("cuda:0")
size = [1, 1, 24, 24, 24]
cfg = stage_cfgs(3, size, vector=[3, 3, 3])["morph"]
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for seed in (5, 5, 6, 7, 8):
    torch.manual_seed(seed)
    x = torch.rand(*size, device=dev)
    conv = torch.nn.Conv3d(1, 3, 3, 1, 1).eval().to(dev)
    probe = AdvMorph(3, dict(cfg), device=dev)
    probe.init_parameters()
    v0 = probe.param.detach().clone()
    n2 = float(_ops.morph_unorm2(v0, size, probe._morph_cfg(), 1.0).item()) ** 0.5
    eps = 0.40 * 256.0 / n2
    outs, info = [], []
    for graph in (False, True, True):
        c = dict(cfg)
        c["epsilon"] = eps
        t = AdvMorph(3, c, device=dev)
        sol = ComposeAdversarialTransformSolver([t], divergence_types=["mse", "contour"], divergence_weights=[1.0, 0.5],
                                                if_norm_image=True, min_intensity=0.0, max_intensity=1.0)
        sol.use_cuda_graph = graph
        sol.graph_capture_after = 0
        sol.spec_first_ratio = 1.0
        init = sol.get_init_output(conv, x)
        t.init_parameters()
        t.param = v0.clone()
        counts = []
        if not graph:
            orig = t._nb_steps

            def spy(orig=orig, counts=counts, t=t):
                n = orig()
                nn = float(_ops.morph_unorm2(t.param.detach(), t.data_size, t._morph_cfg(), t._scale()).item()) ** 0.5
                if not counts or counts[-1][1] != round(nn, 3):
                    counts.append((n, round(nn, 3), round(nn / 2.0 ** n, 4)))
                return n
            t._nb_steps = spy
        sol.optimizing_transform(model=conv, data=x, init_output=init, optimize_flags=[True], n_iter=n_iter,
                                 step_sizes=[1.0])
        outs.append(t.param.detach().clone())
        info.append((counts, getattr(sol, "spec_trace", None), getattr(sol, "graph_iter_redos", 0),
                     getattr(sol, "graph_redos", 0), getattr(sol, "graph_replays", 0),
                     sorted(k for v in sol._graphs.values() if isinstance(v, dict) and v["refs"][1]() is t for k in v["graphs"])))
    d1 = float((outs[1] - outs[0]).norm() / outs[0].norm())
    d2 = float((outs[2] - outs[0]).norm() / outs[0].norm())
    d12 = float((outs[2] - outs[1]).norm() / outs[1].norm())
    print("seed %d: graph vs eager %.2e, second graph run vs eager %.2e, graph vs graph %.2e" % (seed, d1, d2, d12))
    print("   eager (count, |u|, |u|/2^n):", info[0][0])
    for k in (1, 2):
        print("   graph run %d: trace %s iter_redos %d loop_redos %d replays %d graphs %s" % ((k,) + info[k][1:]))
