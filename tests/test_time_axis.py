"""The time axis: sector light effects and map-object animation stepped tic by tic, as the
reference's thinkers do (src/thinkers.rs, src/lights.rs, src/map_objects.rs:63-95, src/game.rs:456-482).

The reference seeds its effects from rand::thread_rng(), so no two of its runs agree; the oracle and the product both
define the random draws as one PCG32 stream per seed, consumed in the reference's thinker order (documented in
oracle/drr_oracle.cpp: Thinkers and include/drr.h: drr_scene_set_tic).  Checked here: the two independent implementations
agree on the world of every (tic, seed); the deterministic effects follow lights.rs by hand; frames at a later tic match.
"""
from __future__ import annotations

import numpy as np
import pytest

import common
from common import drr, orc, synth_wad

KIND = "e1m1_time"  # the E1M1-class map with light-effect sectors and the animated things' extra sprite frames


def test_world_state_agrees_with_oracle_over_tics_and_seeds():
    path, gm = common.wad(KIND)
    scene = drr.Scene(path, "E1M1", 160, 100)
    game = orc.Game(path, "E1M1", 160, 100)
    l0, o0 = scene.world_state()
    changed_lights = changed_objs = 0
    for seed in (0, 7, 0xD00D1993):
        for tic in (0, 1, 2, 3, 4, 5, 16, 17, 35, 36, 70, 99, 100, 350, 1000):
            scene.set_tic(tic, seed)
            game.set_tic(tic, seed)
            (ls, os_), (lo, oo) = scene.world_state(), game.world_state()
            assert (ls == lo).all(), (seed, tic)
            assert (os_ == oo).all(), (seed, tic)
            changed_lights += int((ls != l0).sum())
            changed_objs += int((os_ != o0).any(1).sum())
    assert changed_lights > 100 and changed_objs > 100, "nothing moves: the test would prove nothing"
    scene.set_tic(0, 5)  # tic 0 is the WAD as loaded, whatever the seed
    ls, os_ = scene.world_state()
    assert (ls == l0).all() and (os_ == o0).all()
    # different seeds give different worlds (the random effects), the same seed the same world
    scene.set_tic(200, 1)
    a = scene.world_state()[0].copy()
    scene.set_tic(200, 2)
    b = scene.world_state()[0].copy()
    scene.set_tic(200, 1)
    assert (scene.world_state()[0] == a).all() and (a != b).any()


def test_deterministic_light_effects_follow_lights_rs():
    """Glow (special 8) and the synchronised strobes (12, 13) use no random number: their levels follow from lights.rs by hand."""
    path, gm = common.wad(KIND)
    scene = drr.Scene(path, "E1M1", 160, 100)
    specials = np.array([s.special for s in gm.sectors])
    lights0 = np.array([s.light for s in gm.sectors])
    assert len(specials) == len(scene.world_state()[0])
    hist = []
    for tic in range(0, 120):
        scene.set_tic(tic, 3)
        hist.append(scene.world_state()[0].copy())
    hist = np.array(hist)  # [tic][sector]
    for k in np.nonzero(specials == 8)[0]:  # lights.rs:192-212: down 8 per tic until <= min, then up 8 per tic until >= max
        lvl = hist[:, k].astype(int)
        steps = np.diff(lvl)
        assert set(np.unique(steps)) <= {-8, 0, 8}, k
        assert lvl[0] == lights0[k] and lvl.max() <= lights0[k], k
        down = np.nonzero(steps == -8)[0]
        if len(down):  # it starts by going down, one GLOW_SPEED per tic
            assert down[0] == 0 and (steps[: np.argmax(steps != -8) or len(steps)] == -8).all(), k
    for sp, dark in ((12, 35), (13, 15)):  # lights.rs:144-164: count 1 -> dark at tic 1 for dark_time tics, bright for 5
        for k in np.nonzero(specials == sp)[0]:
            lvl = hist[:, k]
            hi = lights0[k]
            assert lvl[0] == hi and lvl[1] != hi, (sp, k)
            lo = lvl[1]
            period = dark + 5
            for tic in range(1, 120):
                want = lo if (tic - 1) % period < dark else hi
                assert lvl[tic] == want, (sp, k, tic)


@pytest.mark.parametrize("tic,seed", [(17, 7), (100, 1)])
def test_front_end_replay_equals_oracle_at_a_later_tic(tic, seed):
    """Front-end + column binning (host restatement of the bin kernel) replayed on the CPU == the oracle's frame, at tic > 0."""
    W, H, n = 160, 100, 5
    path, gm = common.wad(KIND)
    game = orc.Game(path, "E1M1", W, H)
    game.set_tic(tic, seed)
    ts = tic / 35.0
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 256)[::37], n)
    ctx = drr.Context(W, H, 0, len(views), _host_only=True)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.upload_assets(ctx)
    scene.set_tic(tic, seed)
    assert scene.emit_views(ctx, views, timestamp=ts) == []
    lit = 0
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]), timestamp=ts)
        got = common.replay_binned_frame(ctx, k)
        assert (got == ref).all(), (k, int((got != ref).any(2).sum()))
        game.set_tic(0, seed)
        lit += int((game.render(float(v[0]), float(v[1]), float(v[2]), timestamp=ts) != ref).any())
        game.set_tic(tic, seed)
    assert lit > 0, "the frames do not depend on the tic: the test would prove nothing"
    # and the device front-end's code (run on the CPU) records the same lists as the host front-end at this tic
    b = drr.Context(W, H, 0, len(views), _host_only=True)
    scene.upload_assets(b)
    assert scene.emit_views_device(b, views, timestamp=ts, _on_host=True) == []
    for which, dt in ((1, drr.SEG_DTYPE), (2, drr.PLANE_DTYPE), (7, drr.COL_DTYPE), (8, np.uint32), (9, np.uint32)):
        assert ctx._list(which, dt).tobytes() == b._list(which, dt).tobytes(), which


@pytest.mark.gpu
@pytest.mark.parametrize("tic,seed", [(35, 3), (211, 9)])
def test_gpu_frames_match_oracle_at_a_later_tic(tic, seed):
    """Host front-end and device front-end, both at tic > 0, through the bin and tile kernels == the oracle's frames."""
    W, H, n = 320, 200, 12
    path, gm = common.wad(KIND)
    game = orc.Game(path, "E1M1", W, H)
    game.set_tic(tic, seed)
    ts = tic / 35.0
    views = common.usable_views(game, synth_wad.walk_viewpoints(gm, 512)[::29], n)
    scene = drr.Scene(path, "E1M1", W, H)
    scene.set_tic(tic, seed)
    ctx = drr.Context(W, H, 0, n)
    scene.upload_assets(ctx)
    assert scene.emit_views(ctx, views, timestamp=ts) == []
    ctx.submit()
    crcs = ctx.read_checksums(0, n)
    for k, v in enumerate(views):
        ref = game.render(float(v[0]), float(v[1]), float(v[2]), timestamp=ts)
        assert (ctx.read_framebuffer(k) == ref).all(), k
        assert int(crcs[k]) == drr.checksum_numpy(ref)
    ctx.reset()
    assert scene.emit_views_device(ctx, views, timestamp=ts) == []
    ctx.draw()
    assert ctx.read_checksums(0, n).tolist() == crcs.tolist()
    scene.set_tic(tic + 40, seed)  # the device front-end's map tables follow the world
    ctx.reset()
    assert scene.emit_views_device(ctx, views, timestamp=ts) == []
    ctx.draw()
    assert ctx.read_checksums(0, n).tolist() != crcs.tolist()
