"""CPU: host-side product logic (sampling tables, panorama lattice, module bookkeeping) against the oracle and the
golden fixtures."""

import numpy as np
import torch

import cases as K
import spgan_oracle as O
from spgan_b200 import generator, grids, panorama
from spgan_b200.models import ops, spgan_ops, spgan_ops_gs, spherenet


def test_product_grids_bit_exact_vs_golden():
    g = K.load("grids.npz")
    for name, c in K.load_json("grid_cases.json").items():
        grid = grids.sampling_grid(c["h"], c["h"], c["cp"])
        assert grid.dtype == np.float32 and grid.shape == (1, 3 * c["h"], 3 * c["h"], 2)
        assert np.array_equal(grid.view(np.uint32), g[name].view(np.uint32)), name


def test_product_grids_bit_exact_vs_oracle_sweep():
    for x_st in (0, 4, 9):
        for y_st in range(0, 140, 11):
            cp = K.train_cp(x_st, y_st)
            for h in (35, 29, 23, 17, 53):
                a = grids.sampling_grid(h, h, cp)
                b = O.gen_sampling_grid(h, h, cp)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (x_st, y_st, h)


def test_dense_pattern_matches_oracle():
    cp = K.train_cp(3, 130)
    assert np.array_equal(grids.sampling_pattern_dense(17, 17, cp), O.create_sampling_pattern(17, 17, cp))


def test_grid_cache_hits_and_batches():
    cache = grids.GridCache(max_entries=4)
    cps = [K.train_cp(1, 2), K.train_cp(3, 4)]
    g = cache.batch(17, 17, cps, 2, "cpu")
    assert g.shape == (2, 51, 51, 2) and cache.misses == 2
    cache.batch(17, 17, cps, 2, "cpu")
    assert cache.hits == 2
    one = cache.batch(17, 17, K.test_cp(2, 7, 27), 5, "cpu")
    assert one.shape[0] == 1  # test mode: one grid shared by the batch
    for i in range(6):
        cache.get(11, 11, K.train_cp(i, i), "cpu")
    assert len(cache._store) == 4


def test_panorama_plan_and_cursors_match_oracle_and_golden():
    ref = K.load_json("lattice.json")
    for key, hw in (("384x768", (384, 768)), ("768x1536", (768, 1536))):
        pl = panorama.plan(*hw)
        for k, v in ref[key].items():
            assert pl[k] == v, (key, k)
        po = O.close_loop_plan(*hw)
        for it, (ix, iy) in enumerate(panorama.positions(pl)):
            a = panorama.patch_inputs(pl, ix, iy, it, pl["lat_h"], pl["lat_w"])
            b = O.patch_coords_partial(po, ix, iy, po["lat_h"], po["lat_w"], it)
            assert a == b
    assert len(panorama.positions(panorama.plan(384, 768))) == 60
    assert len(panorama.positions(panorama.plan(768, 1536))) == 180


def test_circular_slice_and_assign_match_oracle():
    t = torch.arange(2 * 3 * 7 * 10, dtype=torch.float32).view(2, 3, 7, 10)
    for (ys, ye) in ((0, 4), (7, 12), (10, 14), (12, 16), (21, 25)):
        a = panorama.circular_slice(t, 10, 1, 5, ys, ye)
        b = O.circular_slice(t, 10, 1, 5, ys, ye)
        assert torch.equal(a, b)
        m1, m2 = torch.zeros_like(t), torch.zeros_like(t)
        v = torch.randn(2, 3, 4, ye - ys)
        panorama.circular_assign(m1, 10, 1, 5, ys, ye, v)
        O.circular_assign(m2, 10, 1, 5, ys, ye, v)
        assert torch.equal(m1, m2)


def test_generator_state_dict_matches_reference_manifest():
    g = generator.Generator()
    manifest = K.load_json("generator_manifest.json")
    sd = g.state_dict()
    assert set(sd) == set(manifest)
    for k, shape in manifest.items():
        assert list(sd[k].shape) == shape, k
    g.load_state_dict(K.generator_state_dict())  # strict


def test_spatial_size_bookkeeping():
    ts = generator.Generator().texture_synthesizer
    assert ts.calc_out_spatial_size(11, return_list=True) == [19, 17, 31, 29, 55, 53, 103, 101]
    assert ts.calc_in_spatial_size(101, return_list=True) == O.ts_in_sizes(101)
    assert generator.Generator().structure_synthesizer.calc_out_spatial_size(35) == 11


def test_spherical_conv_initialises_to_centre_delta():
    cfg = generator.default_config()
    m = spgan_ops_gs.ModulatedConv2d(7, 4, 3, 8, no_zero_pad=True, config=cfg, side="ss", deal_coords=True)
    w = m.weight.detach()
    assert w.shape == (1, 4, 7, 3, 3)
    assert torch.all(w[..., 1, 1] == 1) and w.sum() == 4 * 7
    s = spherenet.SphereConvBatchDiffFixBorderGNoGrad(3, 3)
    assert s.weight.shape == (3, 3, 3, 3) and s.weight.sum() == 9 and abs(s.scale - 1 / np.sqrt(27)) < 1e-12


def test_module_surface_matches_reference_names():
    for mod, names in ((ops, ["PixelNorm", "make_kernel", "Upsample", "Downsample", "Blur", "EqualConv2d", "EqualLinear",
                             "ScaledLeakyReLU", "ModulatedConv2d", "NoiseInjection", "ConstantInput", "StyledConv", "ToRGB"]),
                       (spgan_ops, ["ToRGB", "Upsample", "ModulatedConv2d", "SphereModulatedConv2d", "StyledConv"]),
                       (spgan_ops_gs, ["ModulatedConv2d", "StyledConv", "Blur", "EqualConv2d", "EqualLinear"]),
                       (spherenet, ["GridGenerator", "GridSamplerNewTextureNoGrad", "GridGeneratorPatchCoordsFixBorder",
                                    "GridSamplerNewTexture", "SphereConv2d", "IncreIntervalSphereConv2d",
                                    "SphereConvBatchDiffFixBorderGNoGrad"])):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)


def test_grid_factorisation_equals_dense_grid_bitwise():
    """The row-factor / column-factor decomposition used by the device-side grid assembly reproduces the dense grid
    (and therefore the golden grids of the reference) bit for bit, for training and test-mode windows."""
    import numpy as np
    import cases as K
    from spgan_b200 import grids
    rng = np.random.RandomState(3)
    cps = [K.train_cp(int(rng.randint(0, 10)), int(rng.randint(0, 140)), 35) for _ in range(12)]
    cps += [K.test_cp(2, 7, 27), K.test_cp(0, 9, 3), K.train_cp(9, 139, 35), K.train_cp(0, 0, 35)]
    for cp in cps:
        for h in (35, 29, 23, 17, 53):
            want = grids.sampling_grid(h, h, cp)
            got = grids.assemble_reference(h, h, cp)
            assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), (h, cp)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the arm the driver runs first, CPU only) prints one JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample-patches", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0
    # the real reference when it is present (/root/reference here, oracle/_ref on the GPU box), else the oracle port
    sys.path.insert(0, os.path.join(root, "oracle"))
    import refimport
    assert line["cpu_baseline"]["kind"] == ("reference" if refimport.available() else "port")
    assert len(out.stdout.strip().splitlines()) == 1, "the arm must print exactly one line on stdout"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # other ranks of a torchrun launch exit 0 without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference"], capture_output=True, text=True,
                         timeout=120, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_position_group_channel_map_is_block_diagonal():
    """Several lattice positions per generator call (grids.PositionGroup): the reference builds its flat (1,B*C)++(1,B*3)
    concat table per call, so the table of a stacked batch is one copy of the per-call table per position, with the source
    sample offset by the position's first row (models/spgan_ops_gs.py:792-814)."""
    import numpy as np
    import spgan_b200.functional as SF
    Bg, G, C, nc, Cp = 4, 3, 256, 3, 320
    one = SF._sphere_chan_map(Bg, C, nc, Cp, True, "cpu").numpy().view(np.uint32)
    full = SF._sphere_chan_map(Bg * G, C, nc, Cp, True, "cpu", group=Bg).numpy().view(np.uint32)
    assert full.shape == (Bg * G, Cp)
    for i in range(G):
        blk = full[i * Bg:(i + 1) * Bg]
        valid = one != 0xFFFFFFFF
        assert np.array_equal(blk != 0xFFFFFFFF, valid)
        assert np.array_equal(blk[valid] & 0x80007FFF, one[valid] & 0x80007FFF)            # kind and source channel
        assert np.array_equal((blk[valid] >> 15) & 0xFFFF, ((one[valid] >> 15) & 0xFFFF) + i * Bg)  # source sample
    # the oracle's statement of the same quirk for one call: group g reads flat channels [g*Ct, (g+1)*Ct)
    Ct = C + nc
    for g in range(Bg):
        for k in (0, 1, 255, 256, 258):
            flat = g * Ct + k
            m = int(one[g, k])
            if flat < Bg * C:
                assert (m >> 31) == 0 and ((m >> 15) & 0xFFFF) == flat // C and (m & 0x7FFF) == flat % C
            else:
                assert (m >> 31) == 1 and ((m >> 15) & 0xFFFF) == (flat - Bg * C) // nc and (m & 0x7FFF) == (flat - Bg * C) % nc


def test_position_group_grids_and_engine_partition():
    import torch
    from spgan_b200 import grids, panorama
    pl = panorama.plan(384, 768)
    cps = [panorama.patch_inputs(pl, ix, iy, it, pl["lat_h"], pl["lat_w"])[0] for it, (ix, iy) in enumerate(panorama.positions(pl)[8:11])]
    cache = grids.GridCache()
    g = cache.group_grid(17, 17, cps, "cpu")
    assert g.shape == (3, 51, 51, 2)
    for i, cp in enumerate(cps):
        assert torch.equal(g[i:i + 1], cache.get(17, 17, cp, "cpu"))
    pg = grids.PositionGroup(cps, 2)
    expanded = cache.batch(17, 17, pg, 6, "cpu")
    assert expanded.shape == (6, 51, 51, 2) and torch.equal(expanded[2], g[1]) and torch.equal(expanded[5], g[2])
    try:
        cache.batch(17, 17, pg, 5, "cpu")
        raise AssertionError("a PositionGroup of 3 x 2 samples must not accept a batch of 5")
    except RuntimeError:
        pass

    class Stub:  # PanoramaEngine only partitions here: no generator call
        pass
    eng = panorama.PanoramaEngine.__new__(panorama.PanoramaEngine)
    pos = [(it, ix, iy) for it, (ix, iy) in enumerate(panorama.positions(pl))][:7]
    group = 3
    items = [list(range(i, min(i + group, len(pos)))) for i in range(0, len(pos), group)]
    assert items == [[0, 1, 2], [3, 4, 5], [6]]
