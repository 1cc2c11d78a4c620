"""Genome / sub-matrix model around the hot path, mirroring chromosight.utils.contacts_map
(contacts_map.py:79-450 HicGenome, 453-638 ContactMap) without the `cooler` package: the
file is read by chromosight_b200.cool.CoolFile and the per-sub-matrix preprocessing
(distance-law detrend, band trimming) runs through chromosight_b200.utils.preprocessing,
i.e. on the GPU.

Only what `detect` / `quantify` need is here: ICE balancing
(`cooler.balance_cooler`), sub-sampling and the dump decorators are out of scope; a file
must carry a `weight` column, or be used raw (`norm="raw"`).
"""
import numpy as np
import pandas as pd

from . import cool as _cool
from .utils import preprocessing as preproc


def chrom_codes(values, names):
    """Index of every chromosome name of `values` in `names` (-1 when absent)."""
    codes, uniques = pd.factorize(pd.Series(values), sort=False)
    look = {str(n): i for i, n in enumerate(names)}
    remap = np.array([look.get(str(u), -1) for u in uniques] + [-1], dtype=np.int64)  # code -1 = missing value
    return remap[codes]


# the device-side fast path of ContactMap.create_mat (tests switch it off to compare both)
FAST_CREATE_MAT = True


class ContactMap:
    """One intra- or inter-chromosomal sub-matrix (contacts_map.py:453-638).  The attributes
    the hot path reads are `matrix`, `detectable_bins`, `max_dist`, `inter`, `name`."""

    def __init__(self, clr, extent, inter=False, detectable_bins=None, max_dist=None,
                 largest_kernel=0, use_norm=True, smooth=False, name=None):
        self.clr = clr
        self.extent = extent
        self.inter = inter
        self.max_dist = max_dist
        self.largest_kernel = largest_kernel
        self.use_norm = use_norm
        self.smooth = smooth
        self.name = name
        self._matrix = None
        self.device_csr = None   # the preprocessed sub-matrix in HBM (fast path of create_mat)
        self.detectable_bins = detectable_bins
        if detectable_bins is None:
            raise ValueError("detectable_bins are required (get_detectable_bins is not mirrored)")

    @property
    def matrix(self):
        """The preprocessed sub-matrix as a scipy matrix (what the reference keeps in
        `ContactMap.matrix`).  After the device-side fast path of create_mat it is copied
        back from HBM on first access only: pattern_detector reads `device_csr` instead."""
        if self._matrix is None and self.device_csr is not None:
            self._matrix = self.device_csr.to_scipy()
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        self._matrix = value

    @property
    def shape(self):
        (s1, e1), (s2, e2) = self.extent
        return (e1 - s1, e2 - s2)

    @property
    def keep_distance(self):
        """Diagonals kept in an intra map: scan distance plus a kernel margin (cm:629-638)."""
        n = self.shape[0]
        return (n if self.max_dist is None else min(self.max_dist, n)) + self.largest_kernel

    def _create_mat_device(self):
        """Fast path of create_mat for balanced maps read by CoolFile: the stored (upper)
        triangle is sliced straight into CSR arrays (no mirrored copy, no sort), uploaded once
        and preprocessed in HBM -- distance-law detrend + clamp for an intra map (cm:607-624;
        the band trim is implicit), median normalisation for an inter map (cm:598-601).  The
        result stays on the device for pattern_detector.  Returns False when the file or the
        options need the general path."""
        if not (self.use_norm and not self.smooth and FAST_CREATE_MAT):
            return False
        (s1, e1), (s2, e2) = self.extent
        if self.inter:
            get = getattr(self.clr, "block_csr", None)
            arrays = get(s1, e1, s2, e2, balance=True) if get else None
            if arrays is None:
                return False
            self.device_csr = preproc.divide_by_median_device(*arrays, shape=(e1 - s1, e2 - s2))
        else:
            get = getattr(self.clr, "upper_band_csr", None)
            keep = self.keep_distance
            arrays = get(s1, e1, keep, balance=True) if get else None
            if arrays is None:
                return False
            self.device_csr = preproc.detrend_band_device(
                *arrays, n=e1 - s1, detectable_bins=self.detectable_bins[0], max_dist=keep, max_val=10)
        self._matrix = None
        return True

    def create_mat(self):
        """Load, balance, detrend and trim the sub-matrix (cm:527-548)."""
        if self._create_mat_device():
            return
        (s1, e1), (s2, e2) = self.extent
        self.device_csr = None
        self.matrix = self.clr.matrix(sparse=True, balance=self.use_norm)[s1:e1, s2:e2]
        if self.inter:
            # cm:598-601: stored values (NaN -> 0) divided by their median, on the device
            self.matrix.data = preproc.divide_by_median(self.matrix.data)
        else:
            # cm:607-624: distance-law detrend (GPU) then band trim
            self.matrix = preproc.detrend(
                self.matrix, max_dist=self.keep_distance, smooth=self.smooth,
                detectable_bins=self.detectable_bins[0], max_val=10 if self.use_norm else None)
            self.matrix = preproc.diag_trim(self.matrix.tocsr(), self.keep_distance)
        if self.use_norm:
            self.matrix.data[np.isnan(self.matrix.data)] = 0
        else:
            # raw matrices have no NaN: blank the undetectable bins explicitly (cm:541-547)
            m = self.matrix.tolil()
            m[preproc.valid_to_missing(self.detectable_bins[0], m.shape[0]), :] = 0
            m[:, preproc.valid_to_missing(self.detectable_bins[1], m.shape[1])] = 0
            self.matrix = m.tocoo()
        self.matrix.eliminate_zeros()

    def destroy_mat(self):
        self.matrix = None
        self.device_csr = None


class HicGenome:
    """Whole-genome contact map split into sub-matrices (contacts_map.py:79-450)."""

    def __init__(self, path, inter=False, kernel_config=None, smooth=False):
        self.clr = path if isinstance(path, _cool.CoolFile) else _cool.CoolFile(path)
        self.bins = self.clr.bins()[:]
        self.inter = inter
        self.kernel_config = kernel_config
        self.smooth = smooth
        self.sub_mats = None
        self.detectable_bins = None
        self.use_norm = True
        self.max_dist = None
        self.largest_kernel = 3
        if kernel_config is not None:
            self.compute_max_dist()

    def compute_max_dist(self):
        """cm:166-180"""
        try:
            self.max_dist = max(self.kernel_config["max_dist"] // self.clr.binsize, 1)
            self.largest_kernel = max(k.shape[0] for k in self.kernel_config["kernels"])
        except (ValueError, TypeError):
            self.max_dist = None
            self.largest_kernel = 3

    def normalize(self, norm="auto", n_mads=5, threads=1):
        """cm:182-233, minus the balancing itself: existing weights are reused."""
        if norm not in ("auto", "raw", "force"):
            raise ValueError("norm must be one of: auto, raw, force")
        if "weight" not in self.bins.columns or norm == "force":
            raise NotImplementedError(
                "ICE balancing (cooler.balance_cooler) is outside this package: balance the "
                "file with cooler first, or use norm='raw' on a file with weights")
        self.use_norm = norm != "raw"
        self.detectable_bins = np.flatnonzero(np.isfinite(self.bins.weight.values))

    def make_sub_matrices(self):
        """Table [chr1, chr2, contact_map] of the intra (and, with inter=True, the upper
        inter) sub-matrices (cm:235-322)."""
        d = self.detectable_bins
        rows = []
        for i1, chr1 in enumerate(self.clr.chromnames):
            for i2, chr2 in enumerate(self.clr.chromnames):
                if not (i1 == i2 or (i1 < i2 and self.inter)):
                    continue
                s1, e1 = self.clr.extent(chr1)
                s2, e2 = self.clr.extent(chr2)
                det = (d[(d >= s1) & (d < e1)] - s1, d[(d >= s2) & (d < e2)] - s2)
                kw = dict(extent=[(s1, e1), (s2, e2)], detectable_bins=det, use_norm=self.use_norm,
                          smooth=self.smooth, name=f"{chr1}-{chr2}")
                if i1 == i2:
                    cm = ContactMap(self.clr, inter=False, max_dist=self.max_dist,
                                    largest_kernel=self.largest_kernel, **kw)
                else:
                    cm = ContactMap(self.clr, inter=True, **kw)
                rows.append({"chr1": chr1, "chr2": chr2, "contact_map": cm})
        self.sub_mats = pd.DataFrame(rows, columns=["chr1", "chr2", "contact_map"])
        return self.sub_mats

    def get_full_mat_pattern(self, chr1, chr2, patterns):
        """Sub-matrix bins -> whole-genome bins (cm:336-364)."""
        full = patterns.copy()
        full["bin1"] = full.bin1 + self.clr.extent(chr1)[0]
        full["bin2"] = full.bin2 + self.clr.extent(chr2)[0]
        return full

    def get_sub_mat_pattern(self, chr1, chr2, patterns):
        """Whole-genome bins -> sub-matrix bins (cm:366-394)."""
        sub = patterns.copy()
        sub["bin1"] = sub.bin1 - self.clr.extent(chr1)[0]
        sub["bin2"] = sub.bin2 - self.clr.extent(chr2)[0]
        return sub

    def bins_to_coords(self, bin_idx):
        """cm:396-414"""
        return self.bins.iloc[np.asarray(bin_idx), :][["chrom", "start", "end"]]

    def coords_to_bins(self, coords):
        """Genomic positions (DataFrame[chrom, pos]) -> whole-genome bin ids, in the order of
        the input, NaN where no bin starts at floor(pos / binsize) * binsize (cm:416-450).
        Fixed-size bins are located by arithmetic; any other table by the reference's join."""
        bs = self.clr.binsize
        if self._uniform_bins():
            codes = chrom_codes(coords.chrom, self._chrom_names)
            k = np.asarray(coords.pos, dtype=np.int64) // bs
            ok = (codes >= 0) & (k >= 0) & (k < self._chrom_nbins[np.maximum(codes, 0)])
            out = np.full(len(k), np.nan)
            out[ok] = self._chrom_first[codes[ok]] + k[ok]
            return out
        pos = (coords.pos // bs) * bs
        key = pd.MultiIndex.from_arrays([self.bins.chrom.astype(str), self.bins.start])
        look = pd.Series(np.arange(len(self.bins)), index=key)
        return look.reindex(pd.MultiIndex.from_arrays([coords.chrom.astype(str), pos])).values

    def _uniform_bins(self):
        """True when every chromosome's bins start at 0, binsize, 2 binsize, ... (cached)."""
        if getattr(self, "_uniform", None) is None:
            self._uniform = False
            bs = self.clr.binsize
            if bs:
                chrom = self.bins.chrom.astype(str).values
                start = self.bins.start.values.astype(np.int64)
                names = [str(c) for c in self.clr.chromnames]
                first = np.array([self.clr.extent(c)[0] for c in names], dtype=np.int64)
                last = np.array([self.clr.extent(c)[1] for c in names], dtype=np.int64)
                owner = np.searchsorted(last, np.arange(len(start)), side="right")
                ok = owner < len(names)
                if ok.all() and (np.asarray(names, dtype=object)[owner] == chrom).all() and \
                        (start == (np.arange(len(start)) - first[owner]) * bs).all():
                    self._uniform = True
                    self._chrom_names, self._chrom_first, self._chrom_nbins = names, first, last - first
        return self._uniform
