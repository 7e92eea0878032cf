"""GPU: the additional bench workloads (bench_workloads.py: BASELINE configs 3, 4, 5, per-level config 2, compaction) at
toy sizes, in-process -- they must keep producing a well-formed JSON line with a roofline and finite numbers, and the
sharding helpers must rebuild a unit identically from its id."""
import argparse
import json

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctx(**kw):
    import bench_workloads as bw
    args = argparse.Namespace(pairs=2, units=0, segments=0, steps=2, warmup=3, no_cpu_baseline=True)
    for k, v in kw.items():
        setattr(args, k, v)
    return bw, bw.Ctx(args, 0, 1, torch.device("cuda", 0), None)


@pytest.mark.parametrize("name,kw", [("c3", dict(units=3, segments=12)), ("c4", dict(units=2, segments=12)),
                                     ("c5", dict(units=2, segments=16)), ("c2levels", dict(pairs=2)),
                                     ("compaction", {})])
def test_workload_runs_and_reports(name, kw, capsys):
    bw, ctx = _ctx(**kw)
    bw.RUNNERS[name](ctx)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["roofline"]["frac"] > 0
    assert line["roofline"]["peak"] > 1000.0
    if name in ("c3", "c4", "c5"):
        assert line["shard_check"]["ranks"] == 1
    if name == "c2levels":
        assert [l["level"] for l in line["levels"]] == [0, 0, 1, 1, 2, 2]
    torch.set_grad_enabled(True)


def test_units_are_rebuilt_identically_from_their_id():
    """the shard check re-runs another rank's units: building a unit twice must give bit-identical inputs and results"""
    bw, ctx = _ctx()
    from super_primitive_b200.solver import AlignmentBatch
    res = []
    for _ in range(2):
        probs = bw.build_pair_units([5, 2], 96, 128, 6, ctx.device, n_geoms=2)
        b = AlignmentBatch(probs)
        b.run_gn(3)
        torch.cuda.synchronize()
        res.append((b.poses.clone(), b.k.clone(), probs[0]['pack'].clone(), probs[1]['trg_rgba'].clone()))
    for x, y in zip(*res):
        assert torch.equal(x, y)
