"""Real-data plumbing (BASELINE.json configs[0]: detect loops on data_test/example.cool).

CPU part: the minimal .cool (HDF5) reader against the tables stored in the fixture, and the
genome / sub-matrix model's bookkeeping.  GPU part: ContactMap.create_mat + pattern_detector
on every chromosome against what the unmodified reference produced from the same file
(tests/golden/make_golden_cool.py), and the sharded detect driver on top."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN

REF_COOL = "/root/reference/data_test/example.cool"


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "cool_example_loops.npz"), allow_pickle=False)


def cool_from_fixture(fx):
    from chromosight_b200.cool import CoolFile
    return CoolFile.from_tables(fx["chrom_names"], fx["chrom_sizes"], fx["bin_chrom"], fx["bin_start"],
                                fx["bin_end"], fx["bin_weight"], fx["pix_bin1"], fx["pix_bin2"],
                                fx["pix_count"], binsize=int(fx["binsize"]))


def loops_config(fx):
    cfg = json.loads(str(fx["config"]))
    cfg["kernels"] = [np.array(fx["kernel"])]
    return cfg


@pytest.mark.skipif(not os.path.exists(REF_COOL), reason="reference test data not present")
def test_cool_reader_reads_the_reference_file(fx):
    from chromosight_b200.cool import CoolFile, load_cool
    c = CoolFile(REF_COOL)
    # what cooler recorded about its own tables
    assert c.info["nbins"] == c.shape[0] == 720 and c.info["nchroms"] == 3
    pix = c.pixels()[:]
    assert c.info["nnz"] == len(pix) and c.info["sum"] == pix["count"].sum()
    assert c.binsize == 1000 and c.chromnames == ["chr1", "chr2", "chr3"]
    assert [c.extent(n) for n in c.chromnames] == [(0, 127), (127, 549), (549, 720)]
    bins = c.bins()[:]
    for a, b in ((bins.start.values, fx["bin_start"]), (bins.end.values, fx["bin_end"]),
                 (pix.bin1_id.values, fx["pix_bin1"]), (pix.bin2_id.values, fx["pix_bin2"]),
                 (pix["count"].values, fx["pix_count"])):
        assert np.array_equal(a, b)
    assert np.array_equal(bins.weight.values, fx["bin_weight"], equal_nan=True)
    mat, chroms, bins2, bs = load_cool(REF_COOL)   # io.py:20-78
    assert mat.shape == (720, 720) and bs == 1000 and (mat.row <= mat.col).all()
    assert list(chroms.start_bin) == [0, 127, 549] and list(chroms.end_bin) == [127, 549, 720]
    with pytest.raises(ValueError):
        CoolFile(__file__)


def test_genome_model_bookkeeping(fx):
    import pandas as pd
    from chromosight_b200.contacts_map import HicGenome
    clr = cool_from_fixture(fx)
    m = clr.matrix(balance=True)[0:127, 0:127]
    assert abs(sp.csr_matrix(np.nan_to_num(m.toarray())) - sp.csr_matrix(np.nan_to_num(m.toarray())).T).max() < 1e-15
    raw = clr.matrix(balance=False)[127:549, 549:720]
    assert raw.shape == (422, 171)
    hg = HicGenome(clr, inter=False, kernel_config=loops_config(fx))
    assert hg.max_dist == int(fx["max_dist"]) and hg.largest_kernel == 17
    hg.normalize()
    assert len(hg.detectable_bins) == int(np.isfinite(fx["bin_weight"]).sum())
    assert len(hg.make_sub_matrices()) == 3                      # test_contacts_map.py:50-65
    hg2 = HicGenome(clr, inter=True, kernel_config=loops_config(fx))
    hg2.normalize()
    assert len(hg2.make_sub_matrices()) == 6
    # test_contacts_map.py:130-136: chr1:0 and chr2:4000 are bins 0 and 131
    idx = hg.coords_to_bins(pd.DataFrame({"chrom": ["chr1", "chr2"], "pos": [0, 4000]}))
    assert list(idx) == [0, 131]
    assert list(hg.bins_to_coords([131]).start) == [4000]
    t = pd.DataFrame({"bin1": [3], "bin2": [5]})
    assert hg.get_sub_mat_pattern("chr2", "chr2", hg.get_full_mat_pattern("chr2", "chr2", t)).equals(t)
    with pytest.raises(NotImplementedError):
        hg.normalize(norm="force")


def test_direct_csr_extraction_matches_the_mirrored_matrix():
    """CoolFile.upper_band_csr (the library's fused host pass, cs_band_csr_from_pixels, for every
    count dtype; its numpy twin) and block_csr (inter blocks from the per-file block index)
    against what `clr.matrix(balance=...)[a:b, c:d]` + triu / diag_trim keep (cm:527-624)."""
    from chromosight_b200 import synthetic
    for dtype in (np.float64, np.int32, np.int64):
        clr = synthetic.genome_cool([900, 700, 400, 250], binsize=10_000, n_diags=117, seed=3, inter_density=2e-3)
        clr._pix["count"] = clr._pix["count"].values.astype(dtype)
        off = clr._row_offsets()
        for ch in clr.chromnames:
            s, e = clr.extent(ch)
            for keep in (40, 5000):
                for balance in (True, False):
                    ref = clr.matrix(sparse=True, balance=balance)[s:e, s:e].tocsr()
                    ref = sp.triu(sp.tril(ref, keep)).tocsr()
                    ref.data[np.isnan(ref.data)] = 0
                    ref.eliminate_zeros()
                    nat = clr._band_csr_native(int(off[s]), int(off[e]), s, e, keep, balance)
                    assert nat is not None
                    for indptr, idx, v in (nat, clr.upper_band_csr(s, e, keep, balance=balance)):
                        m = sp.csr_matrix((v, idx, indptr), shape=(e - s, e - s))
                        assert m.has_canonical_format and m.nnz == ref.nnz and abs(m - ref).max() == 0
        names = clr.chromnames
        for i in range(len(names)):
            for j in range(i + 1, len(names)):
                (s1, e1), (s2, e2) = clr.extent(names[i]), clr.extent(names[j])
                indptr, idx, v = clr.block_csr(s1, e1, s2, e2)
                m = sp.csr_matrix((v, idx, indptr), shape=(e1 - s1, e2 - s2))
                a, b = m.toarray(), clr.matrix(sparse=True, balance=True)[s1:e1, s2:e2].toarray()
                assert m.has_canonical_format
                assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
        assert clr.block_csr(*clr.extent(names[1]), *clr.extent(names[0])) is None   # below the diagonal
        # the library's host helpers against their numpy twins
        assert clr._native() is not None and clr._lex_sorted()
        order, starts = clr._inter_index()
        clr._inter, clr._lex, clr._nat = None, None, None          # numpy paths
        assert clr._lex_sorted()
        order0, starts0 = clr._inter_index()
        assert np.array_equal(order, order0) and np.array_equal(starts, starts0)
    # an unsorted table is recognised as such (both ways) and falls back to the mirrored matrix
    from chromosight_b200.cool import CoolFile
    pix = clr._pix.iloc[::-1].reset_index(drop=True)
    bins = clr._bins
    for native in (True, False):
        bad = CoolFile.from_tables(clr.chromnames, clr.chromsizes.values, bins.chrom.cat.codes.values,
                                   bins.start.values, bins.end.values, bins.weight.values,
                                   pix.bin1_id.values, pix.bin2_id.values, pix["count"].values, binsize=10_000)
        if not native:
            bad._nat = None
        assert not bad._lex_sorted()
        assert bad.upper_band_csr(0, 900, 40) is None


@pytest.mark.gpu
@pytest.mark.parametrize("fast", [True, False])
def test_example_cool_detect_loops_matches_reference(fx, fast, monkeypatch):
    """fast=True: the sub-matrix is sliced into CSR, detrended and kept in HBM
    (ContactMap._create_mat_device); fast=False: the reference's sequence of host matrices."""
    from chromosight_b200 import contacts_map
    from chromosight_b200.contacts_map import HicGenome
    from chromosight_b200.utils import detection as cud
    monkeypatch.setattr(contacts_map, "FAST_CREATE_MAT", fast)
    cfg = loops_config(fx)
    kernel = cfg["kernels"][0]
    hg = HicGenome(cool_from_fixture(fx), inter=False, kernel_config=cfg)
    hg.normalize()
    hg.make_sub_matrices()
    total = 0
    for _, row in hg.sub_mats.iterrows():
        chrom, cm = row.chr1, row.contact_map
        cm.create_mat()
        assert (cm.device_csr is not None) == fast
        exp = sp.coo_matrix((fx[f"{chrom}_matrix_val"], (fx[f"{chrom}_matrix_row"], fx[f"{chrom}_matrix_col"])),
                            shape=tuple(fx[f"{chrom}_matrix_shape"])).tocsr()
        got = cm.matrix.tocsr()
        assert got.shape == exp.shape and got.nnz == exp.nnz
        assert abs(got - exp).max() <= 1e-12 * abs(exp).max()
        table, windows = cud.pattern_detector(cm, cfg, kernel, full=True)
        n = int(fx[f"{chrom}_n"])
        total += n
        assert len(table) == n
        assert np.array_equal(table.bin1.values, fx[f"{chrom}_bin1"])
        assert np.array_equal(table.bin2.values, fx[f"{chrom}_bin2"])
        assert np.abs(table.score.values - fx[f"{chrom}_score"]).max() <= 1e-5
        lp, lp0 = np.log10(table.pvalue.values), np.log10(fx[f"{chrom}_pvalue"])
        assert np.allclose(lp, lp0, rtol=2e-4, atol=1e-4)
        assert np.allclose(windows, fx[f"{chrom}_windows"], rtol=1e-12, atol=0, equal_nan=True)
        cm.destroy_mat()
    assert total == 135


@pytest.mark.gpu
def test_device_create_mat_equals_host_path_on_a_synthetic_genome(monkeypatch):
    """Intra and inter sub-matrices of a synthetic genome: the device-resident preprocessing
    gives the same matrix (<= 1e-12) and pattern_detector the same table and windows as the
    host sequence of the reference (cm:527-624)."""
    from chromosight_b200 import contacts_map, kernels, synthetic
    from chromosight_b200.contacts_map import HicGenome
    from chromosight_b200.utils import detection as cud
    clr = synthetic.genome_cool([900, 700, 500], binsize=10_000, n_diags=80, seed=5, inter_density=0.02,
                                density_floor=1.0)
    cfg = dict(kernels.loops)
    cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    cfg["max_dist"] = 60 * 10_000
    cfg["max_perc_zero"] = 100.0
    out = {}
    for fast in (True, False):
        monkeypatch.setattr(contacts_map, "FAST_CREATE_MAT", fast)
        hg = HicGenome(clr, inter=True, kernel_config=cfg)
        hg.normalize()
        hg.make_sub_matrices()
        assert len(hg.sub_mats) == 6
        for _, row in hg.sub_mats.iterrows():
            cm = row.contact_map
            cm.create_mat()
            assert (cm.device_csr is not None) == fast
            table, windows = cud.pattern_detector(cm, cfg, cfg["kernels"][0], full=True)
            out[(fast, cm.name)] = (cm.matrix.tocsr().copy(), table, windows)
            cm.destroy_mat()
    n_found = 0
    for (fast, name), (mat, table, windows) in out.items():
        if not fast:
            continue
        mat0, table0, windows0 = out[(False, name)]
        assert mat.shape == mat0.shape and mat.nnz == mat0.nnz
        assert abs(mat - mat0).max() <= 1e-12 * abs(mat0).max()
        assert (table is None) == (table0 is None)
        if table is None:
            continue
        n_found += len(table)
        assert np.array_equal(table.bin1.values, table0.bin1.values)
        assert np.array_equal(table.bin2.values, table0.bin2.values)
        assert np.abs(table.score.values - table0.score.values).max() <= 1e-6
        assert np.allclose(windows, windows0, rtol=1e-12, atol=0, equal_nan=True)
    assert n_found > 0


@pytest.mark.gpu
def test_detect_driver_on_example_cool(fx):
    from chromosight_b200 import driver
    from chromosight_b200.contacts_map import HicGenome
    cfg = loops_config(fx)
    hg = HicGenome(cool_from_fixture(fx), inter=False, kernel_config=cfg)
    hg.normalize()
    table, windows = driver.detect(hg, cfg, full=True)
    assert list(table.columns) == ["chrom1", "start1", "end1", "chrom2", "start2", "end2", "bin1", "bin2",
                                   "kernel_id", "iteration", "score", "pvalue", "qvalue"]
    assert len(table) == len(windows) and 0 < len(table) <= 135
    # global filters of cli:807-848: separated patterns, minimum distance, q-values
    sep = max(cfg["min_separation"] // 1000, 1)
    b = table[["bin1", "bin2"]].values
    d = np.abs(b[:, None, :] - b[None, :, :])
    close = (d[..., 0] < sep) & (d[..., 1] < sep)
    assert close.sum() == len(table)
    assert (np.abs(table.start2 - table.start1) >= cfg["min_dist"]).all()
    assert ((table.qvalue >= table.pvalue - 1e-15) & (table.qvalue <= 1)).all()
    assert (table.chrom1.astype(str) == table.chrom2.astype(str)).all()


@pytest.mark.gpu
def test_quantify_driver_matches_per_chromosome_calls(fx, presets):
    """cmd_quantify plumbing (cli:295-496): positions -> sub-matrix bins -> pattern_detector ->
    back to the input table, best kernel per position, sorted by whole-genome bins."""
    import pandas as pd
    from chromosight_b200 import driver
    from chromosight_b200.contacts_map import HicGenome
    from chromosight_b200.utils import detection as cud
    rng = np.random.default_rng(3)
    clr = cool_from_fixture(fx)
    rows = []
    for chrom in clr.chromnames:
        s, e = clr.extent(chrom)
        n = e - s
        b1 = rng.integers(0, n - 40, size=60)
        b2 = b1 + rng.integers(2, 40, size=60)
        for a, b in zip(b1, b2):
            rows.append((chrom, a * 1000, a * 1000 + 1000, chrom, b * 1000, b * 1000 + 1000))
    rows.append(("chr1", 10 ** 9, 10 ** 9 + 1000, "chr1", 10 ** 9 + 5000, 10 ** 9 + 6000))  # off the map
    bed = pd.DataFrame(rows, columns=["chrom1", "start1", "end1", "chrom2", "start2", "end2"])
    bed = bed.sample(frac=1.0, random_state=1).reset_index(drop=True)
    cfg = dict(presets.borders)
    cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(clr, inter=False, kernel_config=cfg)
    hg.normalize()
    table, windows = driver.quantify(hg, cfg, bed)
    assert len(table) == len(bed) == len(windows)
    assert list(table.columns) == ["chrom1", "start1", "end1", "chrom2", "start2", "end2", "bin1", "bin2",
                                   "score", "pvalue", "qvalue"]
    key = table.bin1.fillna(10 ** 12).values * 10 ** 6 + table.bin2.fillna(0).values
    assert (np.diff(key) >= 0).all()
    # the same positions straight through pattern_detector, per chromosome and kernel
    checked = 0
    for _, row in hg.sub_mats.iterrows():
        s = clr.extent(row.chr1)[0]
        sub = table[(table.chrom1 == row.chr1) & table.bin1.notna()]
        coords = np.stack([sub.bin1.values - s, sub.bin2.values - s], axis=1).astype(np.int64)
        row.contact_map.create_mat()
        best = np.full(len(sub), np.nan)
        for k in cfg["kernels"]:
            t, _ = cud.pattern_detector(row.contact_map, cfg, k, coords=coords, full=True)
            best = np.fmax(best, t.score.values)
        row.contact_map.destroy_mat()
        assert np.allclose(sub.score.values, best, rtol=0, atol=1e-6, equal_nan=True)
        checked += len(sub)
    assert checked == len(bed) - 1
    off = table[table.bin1.isna()]
    assert len(off) == 1 and np.isnan(off.score.values[0]) and np.isnan(off.pvalue.values[0])


def test_driver_host_plumbing(fx):
    """The parts of the sharded drivers that need no GPU: window-count costs per sub-matrix and
    the positions -> sub-matrix bins mapping of cmd_quantify (cli:262-293)."""
    import pandas as pd
    from chromosight_b200 import driver, sharding
    from chromosight_b200.contacts_map import HicGenome
    cfg = loops_config(fx)
    hg = HicGenome(cool_from_fixture(fx), inter=True, kernel_config=cfg)
    hg.normalize()
    hg.make_sub_matrices()
    costs = driver.unit_costs(hg)
    D = int(fx["max_dist"])
    sizes = {"chr1": 127, "chr2": 422, "chr3": 171}
    exp = []
    for _, row in hg.sub_mats.iterrows():
        a, b = sizes[row.chr1], sizes[row.chr2]
        exp.append(float(a * b) if row.chr1 != row.chr2 else float((min(D, a) + 1) * a))
    assert costs == exp
    parts = sharding.partition_units(costs, 2)
    assert sorted(parts[0] + parts[1]) == list(range(6))
    pos = pd.DataFrame({"chrom1": ["chr2", "chr1", "chr2", "chr9"], "pos1": [4500, 0, 10 ** 9, 5],
                        "chrom2": ["chr2", "chr1", "chr2", "chr9"], "pos2": [9999, 126999, 10 ** 9, 9]})
    idx, coords = driver._chrom_positions(driver._locate_positions(pos, hg), hg, "chr2", "chr2")
    assert list(idx) == [0] and coords.tolist() == [[4, 9]]       # the off-map position is dropped
    idx, coords = driver._chrom_positions(driver._locate_positions(pos, hg), hg, "chr1", "chr1")
    assert list(idx) == [1] and coords.tolist() == [[0, 126]]
    idx, coords = driver._chrom_positions(driver._locate_positions(pos, hg), hg, "chr3", "chr3")
    assert len(idx) == 0 and coords.shape == (0, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("case,expected_n", [("loops_default", 89), ("loops_nb", 59), ("borders", 57),
                                             ("hairpins", 55)])
def test_detect_driver_reference_held_answers(fx, presets, case, expected_n):
    """SURVEY 8f-4 against known answers the reference itself holds: `chromosight test` reports
    "89 patterns detected" (cli:185-199) and docs/notebooks/detect/example_{loops,borders,
    hairpins}.tsv list 59 / 57 / 55 patterns (plot_output.ipynb:13-15).  The fixture
    (tests/golden/make_golden_cli.py) is the same chain driven through the unmodified reference's
    functions; it reproduces those four answers exactly, coordinates and scores included.  The
    sharded driver must return the same table: same patterns in the same order, scores to 1e-5,
    p- and q-values to 1e-4 relative in log10."""
    from chromosight_b200 import driver
    from chromosight_b200.contacts_map import HicGenome
    z = np.load(os.path.join(GOLDEN, "cli_example.npz"), allow_pickle=False)
    cfg = dict(getattr(presets, str(z[f"{case}_pattern"])))
    cfg.update(eval(str(z[f"{case}_override"]), {"__builtins__": {}}))
    cfg["kernels"] = [np.array(k) for k in cfg["kernels"]]
    hg = HicGenome(cool_from_fixture(fx), inter=False, kernel_config=cfg)
    hg.normalize()
    table, windows = driver.detect(hg, cfg, full=True)
    assert len(table) == expected_n == len(z[f"{case}_bin1"]) == len(windows)
    assert np.array_equal(table.bin1.values, z[f"{case}_bin1"])
    assert np.array_equal(table.bin2.values, z[f"{case}_bin2"])
    assert np.array_equal(table.kernel_id.values, z[f"{case}_kernel_id"])
    assert np.array_equal(table.iteration.values, z[f"{case}_iteration"])
    assert np.abs(table.score.values - z[f"{case}_score"]).max() <= 1e-5
    for col in ("pvalue", "qvalue"):
        a, b = np.log10(table[col].values), np.log10(z[f"{case}_{col}"])
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        assert np.allclose(a[fin], b[fin], rtol=2e-4, atol=1e-4)
    if f"{case}_held_bin1" in z.files:
        # the TSV in the reference's repository: identical pattern coordinates
        held = set(zip(z[f"{case}_held_bin1"].tolist(), z[f"{case}_held_bin2"].tolist()))
        assert held == set(zip(table.bin1.tolist(), table.bin2.tolist()))
