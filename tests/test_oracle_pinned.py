"""Pins oracle/task_oracle.py (numpy restatement) to the reference: (1) against the committed golden
vectors produced by the unmodified reference code, (2) live against the reference when /root/reference
is mounted (build container only)."""
import numpy as np
import pytest

from oracle import task_oracle as O
from tests.golden_util import COMPARE, SCENARIOS, Golden, assert_field, load_assets


def oracle_from_golden(g):
    tables, mocap, obs_norm = load_assets()
    s, c = O.new_state(g.N, mocap, obs_norm, g.init["total_mass"], tables.dof_lower, tables.dof_upper, O.Params())
    for k, v in g.init.items():
        if k in s:
            s[k] = v.copy().reshape(s[k].shape).astype(s[k].dtype)
    return s, c


def inject_sim(outs, seen=None):
    it = iter(outs)

    def simulate(s, tau, ext):
        if seen is not None:  # what the oracle hands to the simulator
            seen["tau"].append(tau.copy())
            seen["ext"].append(None if ext is None else ext.copy())
        o = next(it)
        for k, v in o.items():
            s[k] = v.copy()
    return simulate


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_matches_golden(name):
    g = Golden(name)
    s, c = oracle_from_golden(g)
    saw_reset = False
    for t, st in enumerate(g.step):
        seen = {"tau": [], "ext": []}
        ids = O.step(s, c, st["actions"], st["noise"], inject_sim(st["sim"], seen))
        assert np.array_equal(ids, st["env_ids"]), f"step {t}: compacted reset ids differ"
        # the simulator inputs (north star: PD torques within 1e-5): T:520 per substep, T:498-502 push on substep 0 only
        for j in range(2):
            assert_field(f"tau{j}", seen["tau"][j], st["tau"][j], "float", ctx=f"{name} step {t} ")
        assert_field("push", seen["ext"][0], st["push"], "float", ctx=f"{name} step {t} ")
        assert seen["ext"][1] is None
        saw_reset |= len(ids) > 0
        for k, kind in COMPARE.items():
            assert_field(k, s[k], st["after"][k], kind, ctx=f"{name} step {t} ")
        assert_field("stacked_rewards", s["stacked_rewards"], st["stacked_rewards"], "float", ctx=f"{name} step {t} ")
    if name != "walk_perturb":
        assert saw_reset, "scenario should exercise reset_idx"


def test_golden_perturb_scenario_pushes():
    g = Golden("walk_perturb")
    assert any(np.abs(st["push"]).max() > 100 for st in g.step), "the golden must hold non-zero pelvis pushes"
    assert any(st["after"]["pert_on"].any() for st in g.step)
    assert any((st["after"]["magnitude"] > 0).any() for st in g.step)


def test_known_answers_from_assets():
    tables, mocap, obs_norm = load_assets()
    assert tables.num_bodies == 38 and tables.num_dofs == 33 and tables.num_links == 34
    assert abs(tables.total_mass() - 104.48712) < 1e-9  # SURVEY section 4; T:917 hard-codes 104.48
    assert mocap.shape == (3600, 36) and obs_norm.shape == (2, 37)
    assert np.allclose(np.diff(mocap[:, 0]), 0.0005, atol=1e-6)  # T:115
    assert abs(mocap[0, 34] + mocap[0, 35] + 104.48712 * 9.81) < 1.0  # double-support force sum = weight
    assert tables.body_names[O.L_FOOT] == "L_Foot_Link" and tables.body_names[O.R_FOOT] == "R_Foot_Link"
    assert len(tables.pt_link) + len(tables.cyl_link) == 25 * 8 + 36  # 61 primitives: 25 boxes, 36 cylinders


def test_reset_compaction_is_sorted_nonzero():
    rb = np.array([0, 1, 1, 0, 0, 1, 0, 1], np.int64)
    ids, ids32 = O.reset_compact(rb)
    assert ids.tolist() == [1, 2, 5, 7] and ids32.dtype == np.int32
    assert O.reset_compact(np.zeros(5, np.int64))[0].size == 0
