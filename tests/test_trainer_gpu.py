"""GPU: the sync-free TrainStep captured into ONE CUDA graph (generator pass + discriminator pass + both guarded Adam steps) gives
the same parameters as the same step run eagerly -- i.e. nothing in it needs the host (no .item(), no data-dependent control flow)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_whole_train_step_replays_from_one_cuda_graph(cuda):
    from latentsplat_b200.model.types import Prediction
    from latentsplat_b200.runtime import GraphedStep
    from latentsplat_b200.trainer import OptimizerCfg, TrainStep
    from test_trainer_cpu import _Gen, _groups, _setup
    gen0, disc0, batches = _setup(7)
    flat = lambda b: {"context.image": b["context"]["image"].to(cuda), "target.image": b["target"]["image"].to(cuda),
                      "target.near": b["target"]["near"].to(cuda), "target.far": b["target"]["far"].to(cuda)}
    unflat = lambda f: {"context": {"image": f["context.image"]},
                        "target": {"image": f["target.image"], "near": f["target.near"], "far": f["target.far"]}}
    results = []
    for graphed in (False, True):
        gen, disc = copy.deepcopy(gen0).to(cuda), copy.deepcopy(disc0).to(cuda)
        render, combined = _groups()

        def forward_fn(batch, gen=gen):
            r, c = gen(batch["context"]["image"])
            return Prediction(image=r), Prediction(image=c)

        step = TrainStep(forward_fn, gen.parameters(), disc, render, combined, gen.last.weight, OptimizerCfg(lr=1e-2), OptimizerCfg(lr=1e-2))
        fn = lambda f, step=step: step(unflat(f), 5)
        if graphed:
            # capture on weights that are restored afterwards: GraphedStep's warm-up steps must not count
            snap = copy.deepcopy((gen.state_dict(), disc.state_dict()))
            g = GraphedStep(fn, flat(batches[0]), warmup=1)
            gen.load_state_dict(snap[0]); disc.load_state_dict(snap[1])
            for o in (step.g_opt, step.d_opt):                 # fresh Adam state, as in the eager run
                for st in o.opt.state.values():
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.zero_()
            logs = [{k: v.clone() for k, v in g(flat(b)).items()} for b in batches]
        else:
            logs = [fn(flat(b)) for b in batches]
        torch.cuda.synchronize()
        results.append((logs, [p.detach().clone() for p in list(gen.parameters()) + list(disc.parameters())]))
    (le, pe), (lg, pg) = results
    for a, b in zip(le, lg):
        assert torch.allclose(a["generator/total"], b["generator/total"], rtol=1e-4, atol=1e-6)
        assert torch.allclose(a["discriminator/total"], b["discriminator/total"], rtol=1e-4, atol=1e-6)
    for a, b in zip(pe, pg):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)
