"""Generate the golden fixtures under tests/golden/ from the reference tree.

Run ONCE in the build container (``/root/reference`` is not present on the GPU box):

    python tests/golden/make_golden.py

What it writes (all small, committed):

* ``balls7.npz``      -- the reference's ``tests/systems/balls_7_nvt.{gro,xtc}`` decoded to arrays
                         (positions (201,7,3) f32 angstrom, dims (201,6) f32, names, types, ids).
* ``xyz60.npz``       -- ``docs/source/_static/ex_test_files/trajectory.xyz`` (20,60,3) f32 and the
                         doctest known answers (saponify.py:86-87, :174; timesoap.py:50-52,
                         :115-116, :172-173).
* ``ref_lens.npz``    -- the reference's own LENS / coordination-number golden arrays
                         (tests/trajectory/files/{lens,n_c,psi,phi}.npy,
                         tests/systems/lens_2d_{pbc,fbc}.npy, tests/lens/test_lens/c0-c5,
                         tests/neighbors/test_nn/c0-c4) and the psi / phi doctest answers.
* ``ref_soap.npz``    -- the reference's SOAP goldens (tests/soap/test_soap/c0-c5,
                         tests/trajectory/files/soap.npy) at frames 0::10 (21 of 201 frames, to keep
                         the fixture small) and the timeSOAP goldens in full
                         (tests/tsoap/test_tsoap/c0-c2, tests/trajectory/files/tsoap.npy).
* ``coex.npz``        -- the "type O" atoms (2048 of 8192) of the reference's compressed
                         ``tests/systems/coex/test_coex.xtc`` as decoded by the native reader, the
                         SHA-256 of all decoded coordinates, and the reference's spatial-average
                         golden ``tests/analysis/spavg/test_spavg.npy``.
* ``tcorr.npz``       -- the reference's self-TCF goldens (tests/analysis/tcorr/{t_corr,std_dev}.npy)
                         and the doctest value of time_correlations.py:57.
* ``numba_cases.npz`` -- outputs of the reference's numba kernels (imported by path, unmodified)
                         on seeded random inputs: CSR lists for ragged / out-of-box / centres!=env /
                         pbc on-off cases, a 2048-particle fluid LENS block, and timeSOAP on seeded
                         random spectra including zero vectors.  The inputs are regenerated from the
                         seeds by ``tests/golden_cases.py`` so only outputs are stored.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from golden_cases import (  # noqa: E402
    fluid_case,
    random_csr_cases,
    random_soap_spectra,
)

from dynsight_b200.io import read_gro, read_xtc_uncompressed, read_xyz  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_ref_lens, load_ref_timesoap  # noqa: E402

REF_TESTS = REFERENCE_ROOT / "tests"


def main() -> None:
    ref_lens = load_ref_lens()
    ref_ts = load_ref_timesoap()
    assert ref_lens is not None, "reference tree not found"
    assert ref_ts is not None

    # ---- balls_7 fixture ----
    topo, _, _ = read_gro(REF_TESTS / "systems/balls_7_nvt.gro")
    pos, dims = read_xtc_uncompressed(REF_TESTS / "systems/balls_7_nvt.xtc")
    np.savez_compressed(
        HERE / "balls7.npz", positions=pos, dims=dims, names=topo.names, types=topo.types,
        ids=topo.ids,
    )  # fmt: skip

    # ---- doctest trajectory ----
    topo60, pos60 = read_xyz(REFERENCE_ROOT / "docs/source/_static/ex_test_files/trajectory.xyz")
    np.savez_compressed(
        HERE / "xyz60.npz",
        positions=pos60,
        types=topo60.types,
        kat_soap0_sum=8627.847941030795,
        kat_full_len=576,
        kat_norm0_sum=21.987915602216525,
        kat_dist01=0.10292206044570047,
        kat_tsoap_sum=191.3446863900592,
    )

    # ---- the reference's LENS / n_c goldens ----
    d = {
        "trj_lens": np.load(REF_TESTS / "trajectory/files/lens.npy"),
        "trj_n_c": np.load(REF_TESTS / "trajectory/files/n_c.npy"),
        "lens_2d_pbc": np.load(REF_TESTS / "systems/lens_2d_pbc.npy"),
        "lens_2d_fbc": np.load(REF_TESTS / "systems/lens_2d_fbc.npy"),
        # neighbour-list consumers (tests/trajectory/test_trj.py:80,97-99; doctests misc.py:68,187)
        "trj_psi": np.load(REF_TESTS / "trajectory/files/psi.npy"),
        "trj_phi": np.load(REF_TESTS / "trajectory/files/phi.npy"),
        "kat_psi00": np.float64(0.09556616097688675),
        "kat_phi00": np.float64(0.38268664479255676),
    }
    for f in sorted((REF_TESTS / "lens/test_lens").glob("c*.npy")):
        d["lens_" + f.name.split("_")[0]] = np.load(f)
    for f in sorted((REF_TESTS / "neighbors/test_nn").glob("c*.npy")):
        d["nn_" + f.name.split("_")[0]] = np.load(f)
    np.savez_compressed(HERE / "ref_lens.npz", **d)

    # ---- the reference's SOAP / tSOAP goldens ----
    d = {}
    for f in sorted((REF_TESTS / "soap/test_soap").glob("c*.npy")):
        d["soap_" + f.name.split("_")[0]] = np.load(f)[:, ::10]
    d["trj_soap"] = np.load(REF_TESTS / "trajectory/files/soap.npy")[:, ::10]
    for f in sorted((REF_TESTS / "tsoap/test_tsoap").glob("c*.npy")):
        d["tsoap_" + f.name.split("_")[0]] = np.load(f)
    d["trj_tsoap"] = np.load(REF_TESTS / "trajectory/files/tsoap.npy")
    np.savez_compressed(HERE / "ref_soap.npz", **d)

    # ---- numba kernels on seeded random inputs ----
    d = {}
    for k, case in enumerate(random_csr_cases()):
        indptr, indices = ref_lens.neighbor_list_celllist_centers(
            case["pos_env"].astype(np.float64),
            case["pos_cent"].astype(np.float64),
            case["r_cut"],
            case["box"],
            case["pbc"],
        )
        d[f"csr{k}_indptr"] = indptr
        d[f"csr{k}_indices"] = indices
    pos_f, box_f, r_cut_f = fluid_case()
    lens = np.zeros((pos_f.shape[1], pos_f.shape[0] - 1))
    counts = np.zeros((pos_f.shape[1], pos_f.shape[0]))
    prev = None
    for t in range(pos_f.shape[0]):
        p64 = pos_f[t].astype(np.float64)
        cur = ref_lens.neighbor_list_celllist_centers(p64, p64, r_cut_f, box_f, True)
        counts[:, t] = np.diff(cur[0])
        if prev is not None:
            lens[:, t - 1] = ref_lens.lens_from_two_csr(prev[0], prev[1], cur[0], cur[1])
        prev = cur
    d["fluid_lens"] = lens
    d["fluid_counts"] = counts
    d["fluid_last_indptr"] = prev[0]
    d["fluid_last_indices"] = prev[1]
    spectra = random_soap_spectra()
    for delay in (1, 3):
        d[f"tsoap_d{delay}"] = ref_ts.timesoap(spectra, delay=delay)
    d["tsoap_norm"] = ref_ts.normalize_soap(spectra)
    np.savez_compressed(HERE / "numba_cases.npz", **d)
    # ---- compressed XTC + spatial average ----
    import hashlib

    from dynsight_b200.io import read_xtc
    from oracle import spatial_oracle

    topo_c, _, _ = read_gro(REF_TESTS / "systems/coex/test_coex.gro")
    pos_c, dims_c = read_xtc(REF_TESTS / "systems/coex/test_coex.xtc")
    o_index = np.where(topo_c.types == "O")[0]
    spavg = np.load(REF_TESTS / "analysis/spavg/test_spavg.npy")
    # tests/analysis/test_spatialaverage.py:41-61: dataset = x coordinate of the "type O" atoms
    data_c = pos_c[:, o_index, 0].T.astype(np.float64)
    mine = spatial_oracle.spatial_average(pos_c[:, o_index], dims_c, data_c, 5.0)
    print("spatial oracle vs reference golden: max abs diff", np.abs(mine - spavg).max())
    assert np.allclose(mine, spavg, rtol=1e-13, atol=0)
    np.savez_compressed(
        HERE / "coex.npz",
        o_index=o_index.astype(np.int32),
        o_positions=pos_c[:, o_index],
        dims=dims_c,
        spavg=spavg,
        sha256_all=hashlib.sha256(np.ascontiguousarray(pos_c).tobytes()).hexdigest(),
        n_atoms=pos_c.shape[1],
    )
    # ---- self time-correlation function ----
    from oracle import tcf_oracle

    t_corr = np.load(REF_TESTS / "analysis/tcorr/t_corr.npy")
    t_std = np.load(REF_TESTS / "analysis/tcorr/std_dev.npy")
    mine_c, mine_e = tcf_oracle.self_time_correlation(tcf_oracle.reference_walk())
    print("tcf oracle vs reference golden:", np.abs(mine_c - t_corr).max(),
          np.abs(mine_e - t_std).max())
    assert np.allclose(mine_c, t_corr, rtol=1e-12, atol=1e-14)
    assert np.allclose(mine_e, t_std, rtol=1e-12, atol=1e-14)
    np.random.seed(1234)  # doctest of time_correlations.py:44-57
    kat, _ = tcf_oracle.self_time_correlation(np.random.rand(100, 100))
    assert np.isclose(kat[1], 0.005519088806189553, rtol=1e-12)
    c_corr = np.load(REF_TESTS / "analysis/tcorr/c_corr.npy")
    c_std = np.load(REF_TESTS / "analysis/tcorr/ctd_dev.npy")
    cc, ce = tcf_oracle.cross_time_correlation(tcf_oracle.reference_walk())
    print("cross tcf oracle vs reference golden:", np.abs(cc - c_corr).max(), np.abs(ce - c_std).max())
    assert np.allclose(cc, c_corr, rtol=1e-11, atol=1e-13)
    assert np.allclose(ce, c_std, rtol=1e-11, atol=1e-13)
    np.random.seed(1234)  # doctest of time_correlations.py:128-141
    kat_c, _ = tcf_oracle.cross_time_correlation(np.random.rand(100, 100))
    assert np.isclose(kat_c[0], 0.0002474572311281272, rtol=1e-10)
    np.savez_compressed(HERE / "tcorr.npz", t_corr=t_corr, std_dev=t_std, c_corr=c_corr,
                        ctd_dev=c_std, kat_rand_lag1=0.005519088806189553,
                        kat_cross_lag0=0.0002474572311281272)
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
