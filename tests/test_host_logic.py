"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header declares,
handle plumbing, error behaviour without a GPU, wavelength sharding."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "sasktran2_b200.h").read_text()
    names = sorted(set(re.findall(r"\b(sk_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 100
    lib = ctypes.CDLL(str(ROOT / "sasktran2_b200" / "libsasktran2_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_exports_every_symbol_of_the_reference_c_api():
    """Link-completeness: every `sk_*` function of the reference's cpp/include/c_api/*.h (list extracted by
    tests/golden/make_c_api_symbols.py) resolves in libsasktran2_b200.so, so the Rust layer's bindgen prototypes
    (rust/sasktran2-rs/src/bindings/*.rs) link against it in place of libcsasktran2."""
    names = (ROOT / "tests" / "golden" / "c_api_symbols.txt").read_text().split()
    assert len(names) >= 205
    lib = ctypes.CDLL(str(ROOT / "sasktran2_b200" / "libsasktran2_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_off_path_entry_points_fail_with_the_reference_codes():
    lib = ctypes.CDLL(str(ROOT / "sasktran2_b200" / "libsasktran2_b200.so"))
    lib.sk_b200_last_error.restype = ctypes.c_char_p
    lib.sk_geodetic_create.restype = ctypes.c_void_p
    lib.sk_geodetic_create.argtypes = [ctypes.c_double, ctypes.c_double]
    assert lib.sk_geodetic_create(6378137.0, 1 / 298.257) is None
    assert b"Geodetic" in lib.sk_b200_last_error()
    assert lib.sk_engine_calculate_vjp(None, None, None) == -3
    # stored config options round-trip
    lib.sk_config_create.restype = ctypes.c_void_p
    cfg = ctypes.c_void_p(lib.sk_config_create())
    v = ctypes.c_int(-7)
    assert lib.sk_config_get_num_hr_incoming(cfg, ctypes.byref(v)) == 0 and v.value == 110
    assert lib.sk_config_set_stokes_basis(cfg, 2) == 0
    assert lib.sk_config_get_stokes_basis(cfg, ctypes.byref(v)) == 0 and v.value == 2
    n = ctypes.c_int(0)
    assert lib.sk_config_get_num_flux_types(cfg, ctypes.byref(n)) == 0 and n.value == 2
    d = ctypes.c_double(0)
    lib.sk_config_set_successive_orders_damping.argtypes = [ctypes.c_void_p, ctypes.c_double]
    assert lib.sk_config_set_successive_orders_damping(cfg, 0.5) == 0
    assert lib.sk_config_get_successive_orders_damping(cfg, ctypes.byref(d)) == 0 and d.value == 0.5
    lib.sk_config_destroy(cfg)


def test_sk_lapack_dgesv_matches_numpy():
    lib = ctypes.CDLL(str(ROOT / "sasktran2_b200" / "libsasktran2_b200.so"))
    rng = np.random.default_rng(3)
    n, nrhs = 7, 3
    a = np.asfortranarray(rng.standard_normal((n, n)))
    b = np.asfortranarray(rng.standard_normal((n, nrhs)))
    x_ref = np.linalg.solve(a, b)
    ipiv = np.zeros(n, dtype=np.int64)
    ll = ctypes.c_longlong
    lib.sk_lapack_dgesv.restype = ll
    lib.sk_lapack_dgesv.argtypes = [ll, ll, ctypes.c_void_p, ll, ctypes.c_void_p, ctypes.c_void_p, ll]
    assert lib.sk_lapack_dgesv(n, nrhs, a.ctypes.data, n, ipiv.ctypes.data, b.ctypes.data, n) == 0
    np.testing.assert_allclose(b, x_ref, rtol=1e-10, atol=1e-12)
    sing = np.zeros((3, 3), order="F")
    assert lib.sk_lapack_dgesv(3, 1, sing.ctypes.data, 3, ipiv.ctypes.data, b.ctypes.data, n) == 1


def test_atmosphere_revision_and_delta_m_reapply():
    """sk_atmosphere_mark_changed / get_revision (cpp/c_api/atmosphere.cpp:473-507); delta-M scaling can be applied
    again after sk_atmosphere_storage_set_zero + refill, as upstream does on every internal_object()."""
    import sasktran2_b200 as sk
    from sasktran2_b200 import _lib

    cfg = sk.Config()
    cfg.num_streams = 4
    cfg.delta_m_scaling = True
    geo = sk.Geometry1D(0.6, 0.0, 6372000.0, np.linspace(0, 5e4, 6), sk.InterpolationMethod.LinearInterpolation,
                        sk.GeometryType.PlaneParallel)
    atm = sk.Atmosphere(geo, cfg, numwavel=2, num_legendre=8, calculate_derivatives=False)

    def fill():
        atm.storage.ssa[:] = 0.8
        atm.storage.total_extinction[:] = 1e-5
        atm.storage.leg_coeff[:] = (0.7 ** np.arange(8) * (2 * np.arange(8) + 1))[:, None, None]

    fill()
    h = atm.internal_object()
    scaled = atm.storage.total_extinction.copy()
    assert np.all(scaled < 1e-5)
    rev = ctypes.c_ulonglong(99)
    assert _lib.lib().sk_atmosphere_get_revision(h, ctypes.byref(rev)) == 0 and rev.value == 0
    assert _lib.lib().sk_atmosphere_mark_changed(h) == 0
    assert _lib.lib().sk_atmosphere_get_revision(h, ctypes.byref(rev)) == 0 and rev.value == 1
    atm.internal_object()  # untouched: no second scaling
    np.testing.assert_array_equal(atm.storage.total_extinction, scaled)
    fill()                 # refilled without zero_storage(): refused instead of silently mixing scaled / unscaled
    with pytest.raises(sk.SasktranError):
        atm.internal_object()
    atm.zero_storage()
    assert np.all(atm.storage.ssa == 0)
    fill()
    atm.internal_object()
    np.testing.assert_array_equal(atm.storage.total_extinction, scaled)


def test_config_defaults_and_roundtrip():
    import sasktran2_b200 as sk

    c = sk.Config()
    # reference defaults, cpp/lib/config/config.cpp:5-33
    assert c.num_streams == 16 and c.num_stokes == 1 and c.num_threads == 1
    assert c.multiple_scatter_source == sk.MultipleScatterSource.NoSource
    assert c.single_scatter_source == sk.SingleScatterSource.Exact
    assert c.do_backprop is False and c.wf_enabled is True
    c.num_streams = 8
    c.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    assert c.num_streams == 8 and c.multiple_scatter_source == sk.MultipleScatterSource.DiscreteOrdinates


def test_handles_and_mappings_without_gpu():
    import sasktran2_b200 as sk

    cfg = sk.Config()
    geo = sk.Geometry1D(0.6, 0.0, 6372000.0, np.linspace(0, 5e4, 6), sk.InterpolationMethod.LinearInterpolation,
                        sk.GeometryType.PseudoSpherical)
    np.testing.assert_allclose(geo.altitudes(), np.linspace(0, 5e4, 6))
    atm = sk.Atmosphere(geo, cfg, numwavel=3)
    m = atm.storage.get_derivative_mapping("wf_b")
    m2 = atm.storage.get_derivative_mapping("wf_a")
    m2.d_leg_coeff[:] = 1.0
    m.d_ssa[:] = 2.0
    assert atm.storage.derivative_mapping_names == ["wf_a", "wf_b"]  # std::map name order like upstream
    assert m2.is_scattering_derivative and not m.is_scattering_derivative
    again = atm.storage.get_derivative_mapping("wf_b")
    assert np.all(again.d_ssa == 2.0)  # same storage-owned memory
    assert m.num_output == 6
    m.interpolator = np.ones((6, 2))
    assert m.num_output == 2
    atm.internal_object()


def test_engine_create_fails_loudly_without_cuda():
    import sasktran2_b200 as sk
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = sk.Config()
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    geo = sk.Geometry1D(0.6, 0.0, 6372000.0, np.linspace(0, 5e4, 6), sk.InterpolationMethod.LinearInterpolation,
                        sk.GeometryType.PlaneParallel)
    view = sk.ViewingGeometry()
    view.add_ray(sk.GroundViewingSolar(0.6, 0.0, 1.0, 2e5))
    with pytest.raises(sk.SasktranError, match="no CUDA device"):
        sk.Engine(cfg, geo, view)


def test_unsupported_configurations_are_refused():
    import sasktran2_b200 as sk

    geo = sk.Geometry1D(0.6, 0.0, 6372000.0, np.linspace(0, 5e4, 6), sk.InterpolationMethod.LinearInterpolation,
                        sk.GeometryType.PlaneParallel)
    view = sk.ViewingGeometry()
    view.add_ray(sk.GroundViewingSolar(0.6, 0.0, 1.0, 2e5))
    cfg = sk.Config()  # defaults: MS none / SS exact -> not the DO path
    with pytest.raises(sk.SasktranError, match="multiple_scatter_source"):
        sk.Engine(cfg, geo, view)
    cfg.multiple_scatter_source = sk.MultipleScatterSource.DiscreteOrdinates
    with pytest.raises(sk.SasktranError, match="single_scatter_source"):
        sk.Engine(cfg, geo, view)
    cfg.single_scatter_source = sk.SingleScatterSource.DiscreteOrdinates
    cfg.num_stokes = 3
    with pytest.raises(sk.SasktranError, match="num_stokes"):
        sk.Engine(cfg, geo, view)


def test_wavelength_blocks_partition():
    from sasktran2_b200.parallel import wavelength_block

    for n in (0, 1, 7, 8, 100000, 100003):
        for world in (1, 2, 4, 8):
            blocks = [wavelength_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks[:-1], blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _gloo_worker(rank, world, port, nwavel, q):
    import os

    import torch.distributed as dist

    from sasktran2_b200 import scenarios
    from sasktran2_b200.parallel import gather_wavelength_blocks, shard_scenario

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenarios.config1(nwavel=nwavel, nlayers=6)
    shard, start, count = shard_scenario(sc, rank, world)
    # stand-in for the per-rank solve: a deterministic function of the shard's inputs
    local = (shard.total_extinction.sum(axis=0) + 10 * shard.albedo)[:, None, None] * np.ones((1, 2, 1))
    full = gather_wavelength_blocks(local, nwavel, wavelength_axis=0, dst=0)
    wf_local = np.tile(shard.ssa[None, :3, :, None, None], (1, 1, 1, 2, 1))[0]  # [nout=3, nw_local, nlos, 1]
    wf_full = gather_wavelength_blocks(wf_local, nwavel, wavelength_axis=1, dst=0)
    if rank == 0:
        ref = (sc.total_extinction.sum(axis=0) + 10 * sc.albedo)[:, None, None] * np.ones((1, 2, 1))
        q.put((np.array_equal(full, ref), np.array_equal(wf_full[:, :, 0, 0], sc.ssa[:3, :])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    """world_size-2 run of the N>1 path on CPU (gloo): shard by contiguous wavelength blocks, gather on rank 0."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok == (True, True)


def _shared_result_worker(rank, world, port, nwavel, q):
    import os

    import torch.distributed as dist

    import sasktran2_b200 as sk
    from sasktran2_b200 import scenarios
    from sasktran2_b200.parallel import SharedResult, wavelength_block

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = wavelength_block(nwavel, rank, world)
    # every rank builds its block and places it in full-size caller arrays (what bench.py / a sharded host program does)
    blk = scenarios.config2(nwavel=nwavel, nlayers=6, nstr=4, nlos=2, with_wf=True, block=(start, count))
    cfg = sk.Config()
    cfg.num_streams = 4
    geo = sk.Geometry1D(blk.cos_sza, 0.0, blk.earth_radius, blk.altitudes, sk.InterpolationMethod(blk.interp),
                        sk.GeometryType(blk.geotype))
    atm = sk.Atmosphere.from_scenario(blk, geo, cfg, total_wavelengths=nwavel, wavelength_start=start)
    full = scenarios.config2(nwavel=nwavel, nlayers=6, nstr=4, nlos=2, with_wf=True)
    ok_place = (np.array_equal(atm.storage.ssa[:, start:start + count], full.ssa[:, start:start + count]) and
                np.array_equal(atm.storage.get_derivative_mapping("wf_o3_vmr").d_ssa[:, start:start + count],
                               full.mappings["wf_o3_vmr"]["d_ssa"][:, start:start + count]) and
                not atm.storage.ssa[:, :start].any() and not atm.storage.ssa[:, start + count:].any())
    shapes = {"radiance": (nwavel, 2, 1), "wf:wf_o3_vmr": (7, nwavel, 2, 1)}
    shared = SharedResult(shapes, rank, f"test_{port}", dist.barrier, pin=False)
    # stand-in for the per-rank block solve: every rank writes its block of the caller's arrays
    shared.arrays["radiance"][start:start + count] = full.albedo[start:start + count, None, None] + np.arange(start, start + count)[:, None, None]
    shared.arrays["wf:wf_o3_vmr"][:, start:start + count] = full.ssa[:, start:start + count, None, None]
    dist.barrier()
    if rank == 0:
        ok_rad = np.array_equal(shared.arrays["radiance"][:, 0, 0], full.albedo + np.arange(nwavel))
        ok_wf = np.array_equal(shared.arrays["wf:wf_o3_vmr"][:, :, 1, 0], full.ssa)
        q.put((bool(ok_place), bool(ok_rad), bool(ok_wf)))
    else:
        q.put((bool(ok_place),))
    shared.close(dist.barrier)
    dist.destroy_process_group()


def test_two_rank_gloo_shared_result_arrays():
    """world_size-2 run of the sharded end-to-end path's host side on CPU (gloo): block placement in full-size caller
    arrays, result arrays in shared memory that rank 0 (the caller) ends up holding completely."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29741
    procs = [ctx.Process(target=_shared_result_worker, args=(r, 2, port, 13, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(all(g) for g in got), got
