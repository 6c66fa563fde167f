"""Host-side spherical sampling tables.

The reference rebuilds every (1, 3H, 3W, 2) tap grid in numpy float64, per sample, per layer, on every forward
(models/spgan_ops_gs.py:767-781 -> models/spherenet/grid_generator.py:137-283; its `lru_cache` never hits because
the generator object is new each time).  Bit-exact sampling indices are decided by the last ulp of that float64
arithmetic, so the tables stay on the host, in numpy, with the reference's own operation order —
but vectorised over rows, computed once per distinct (size, coords_partial) and kept resident on the device.
"""
from collections import OrderedDict

import numpy as np
import torch

_KEYS = ("p_x_st", "p_x_ed", "p_y_st", "p_y_ed", "circular_flag", "x_total", "y_total", "test_flag", "partial")


def _norm(v):
    """min_max_norm (grid_generator.py:349-352) with start = -1."""
    return (v - np.min(v)) / (np.max(v) - np.min(v)) * 2 + (-1)


def _check_plain(cp):
    if cp.get("full_shape", None) is not None or cp.get("pre_sample_mode", False):
        raise NotImplementedError("coords_partial['full_shape'/'pre_sample_mode'] is not used by spgan.yaml and is not supported")


def lat_range(h, cp):
    """Centre latitudes of the patch rows: grid_generator.py:164-241, plain branch (`full_shape` / `pre_sample_mode`
    are never set by spgan.yaml or the close-loop manager)."""
    _check_plain(cp)
    partial = 0.8  # training hard-codes 0.8 (grid_generator.py:164-167)
    if cp.get("test_flag", False):
        partial = cp.get("partial", partial)
    x_st = cp["p_x_st"] * np.pi * partial
    x_ed = cp["p_x_ed"] * np.pi * partial
    return np.linspace(x_st, x_ed, h) - (np.pi / 2 * partial)


def lon_range(w, cp):
    """Centre longitudes of the patch columns (same source lines)."""
    _check_plain(cp)
    y_st = cp["p_y_st"] * np.pi * 2
    y_ed = cp["p_y_ed"] * np.pi * 2
    if y_ed != 2 * np.pi:
        y_ed = y_ed % (np.pi * 2)
    if cp["circular_flag"]:
        y_ed = y_ed + 2 * np.pi
    return np.linspace(y_st, y_ed, w) - np.pi


def angular_ranges(h, w, cp):
    return lat_range(h, cp), lon_range(w, cp)


def row_factor(h, cp):
    """The part of the sampling pattern that depends only on the ROWS of the window (p_x_st, p_x_ed, x_total, y_total,
    partial): lat_g (h, 3, 3) float64 latitude tap positions in grid units (shared by every column) and lon (h, 3, 3)
    float64 tangent-plane longitude offsets.  createSamplingPattern of GridGeneratorPatchCoordsFixBorder
    (grid_generator.py:137-283), stride 1, 3x3."""
    x_total, y_total = cp["x_total"], cp["y_total"]
    # createKernel (grid_generator.py:303-323)
    d_lat = np.pi / x_total
    d_lon = 2 * np.pi / y_total
    r = np.arange(-1, 2)
    ker_x, ker_y = np.meshgrid(np.tan(r * d_lon), np.tan(r * d_lat) / np.cos(r * d_lon))
    rho = np.sqrt(ker_x ** 2 + ker_y ** 2)
    rho[1][1] = 1e-8
    nu = np.arctan(rho)
    cos_nu, sin_nu = np.cos(nu), np.sin(nu)
    lat_c = lat_range(h, cp)
    t = lat_c[:, None, None]  # rows broadcast against the 3x3 kernel: same elementwise sequence as the row loop (:249-262)
    lat = np.arcsin(cos_nu * np.sin(t) + ker_y * sin_nu * np.cos(t) / rho)
    lon = np.arctan(ker_x * sin_nu / (rho * np.cos(t) * cos_nu - ker_y * np.sin(t) * sin_nu))
    lat_off = lat - lat[:, 1:2, 1:2]                            # get_pattern (:325-335)
    lat_rows = _norm(lat_c)[:, None, None] + lat_off            # add_pattern_to_lat (:337-346)
    lat_g = (lat_rows / 2 + 0.5) * x_total
    return lat_g, lon


def col_factor(w, cp):
    """The part that depends only on the COLUMNS of the window (p_y_st, p_y_ed, circular_flag): the min-max normalised
    column base (w,) float64."""
    return _norm(lon_range(w, cp))


def sampling_pattern(h, w, cp):
    """float64 (lat, lon) tap positions in grid units: lat (h, 3, 3) shared by every column, lon (h, w, 3, 3)."""
    lat_g, lon = row_factor(h, cp)
    lon_cols = lon[:, None, :, :] + col_factor(w, cp)[None, :, None, None]  # (h, w, 3, 3)
    lon_g = (lon_cols / 2 + 0.5) * cp["y_total"]
    return lat_g, lon_g


def sampling_pattern_dense(h, w, cp):
    """The reference's return layout: (1, 3h, 3w, 2) float64, last dim (lat, lon)."""
    lat_g, lon_g = sampling_pattern(h, w, cp)
    out = np.empty((h, 3, w, 3, 2), dtype=np.float64)
    out[..., 0] = lat_g[:, :, None, :]
    out[..., 1] = lon_g.transpose(0, 2, 1, 3)
    return out.reshape(1, 3 * h, 3 * w, 2)


def sampling_grid(h, w, cp):
    """(1, 3h, 3w, 2) float32 grid in F.grid_sample convention (x = lon, y = lat), equal bit for bit to
    ModulatedConv2d.genSamplingPattern (models/spgan_ops_gs.py:410-428) /
    SphereConvBatchDiffFixBorderGNoGrad.genSamplingPattern (models/spherenet/sphere_conv2d.py:147-165)."""
    lat_g, lon_g = sampling_pattern(h, w, cp)
    lat_n = ((lat_g / cp["x_total"]) * 2 - 1).astype(np.float32)
    lon_n = ((lon_g / cp["y_total"]) * 2 - 1).astype(np.float32)
    out = np.empty((h, 3, w, 3, 2), dtype=np.float32)
    out[..., 0] = lon_n.transpose(0, 2, 1, 3)
    out[..., 1] = lat_n[:, :, None, :]
    return out.reshape(1, 3 * h, 3 * w, 2)


def full_sphere_pattern(height, width, kernel_size=(3, 3), stride=(1, 1)):
    """SphereNet's original full-panorama pattern, GridGenerator.createSamplingPattern (grid_generator.py:28-84):
    (1, H*Kh, W*Kw, 2) float64 (lat, lon) pixel positions, longitude wrapped modulo the width."""
    kh, kw = kernel_size
    d_lat = np.pi / height
    d_lon = 2 * np.pi / width
    rx = np.arange(-(kw // 2), kw // 2 + 1)
    if not kw % 2:
        rx = np.delete(rx, kw // 2)
    ry = np.arange(-(kh // 2), kh // 2 + 1)
    if not kh % 2:
        ry = np.delete(ry, kh // 2)
    ker_x, ker_y = np.meshgrid(np.tan(rx * d_lon), np.tan(ry * d_lat) / np.cos(ry * d_lon))
    rho = np.sqrt(ker_x ** 2 + ker_y ** 2)
    if kh % 2 and kw % 2:
        rho[kh // 2][kw // 2] = 1e-8
    nu = np.arctan(rho)
    cos_nu, sin_nu = np.cos(nu), np.sin(nu)
    lat_c = ((np.arange(0, height, stride[0]) / height) - 0.5) * np.pi
    lon_c = ((np.arange(0, width, stride[1]) / width) - 0.5) * (2 * np.pi)
    t = lat_c[:, None, None]
    lat = np.arcsin(cos_nu * np.sin(t) + ker_y * sin_nu * np.cos(t) / rho)
    lon = np.arctan(ker_x * sin_nu / (rho * np.cos(t) * cos_nu - ker_y * np.sin(t) * sin_nu))
    lon = lon[:, None, :, :] + lon_c[None, :, None, None]
    lat = (lat / np.pi + 0.5) * height
    lon = ((lon / (2 * np.pi) + 0.5) * width) % width
    H, W = len(lat_c), len(lon_c)
    out = np.empty((H, kh, W, kw, 2), dtype=np.float64)
    out[..., 0] = lat[:, :, None, :]
    out[..., 1] = lon.transpose(0, 2, 1, 3)
    return out.reshape(1, H * kh, W * kw, 2)


def _tangent_kernel_full(height, width, kernel_size):
    """createKernel of the full-sphere generators (grid_generator.py:86-108, 562-582) and the derived rho / nu terms."""
    kh, kw = kernel_size
    d_lat = np.pi / height
    d_lon = 2 * np.pi / width
    rx = np.arange(-(kw // 2), kw // 2 + 1)
    if not kw % 2:
        rx = np.delete(rx, kw // 2)
    ry = np.arange(-(kh // 2), kh // 2 + 1)
    if not kh % 2:
        ry = np.delete(ry, kh // 2)
    ker_x, ker_y = np.meshgrid(np.tan(rx * d_lon), np.tan(ry * d_lat) / np.cos(ry * d_lon))
    rho = np.sqrt(ker_x ** 2 + ker_y ** 2)
    if kh % 2 and kw % 2:
        rho[kh // 2][kw // 2] = 1e-8
    nu = np.arctan(rho)
    return ker_x, ker_y, rho, np.cos(nu), np.sin(nu)


def _incre_range(n, k, stride):
    """Row / column centres of IncreIntervalGridGenerator (grid_generator.py:464-521): arange(0, n, stride) trimmed by the
    kernel half-width, then re-spread with linspace over [0, n]."""
    if k == 1:
        return np.arange(0, n, stride)
    d = k // 2
    base = np.arange(0, n, stride)
    if stride not in (1, 2):
        raise NotImplementedError
    if k % 2 == 0:
        r = base[d - 1: -d]
    elif stride == 1:
        r = base[d: -d]
    elif d == 1:
        r = base
    else:
        r = base[d - 1: -d + 1]
    return np.linspace(0, n, len(r))


def incre_interval_pattern(height, width, kernel_size=(3, 3), stride=(1, 1), upsample=False):
    """IncreIntervalGridGenerator.createSamplingPattern (grid_generator.py:385-560): (1, H'*Kh, W'*Kw, 2) float64
    (lat, lon) pixel positions; same tangent-plane formulas as the full-sphere pattern on re-spread row / column centres."""
    kh, kw = kernel_size
    sh, sw = stride
    ker_x, ker_y, rho, cos_nu, sin_nu = _tangent_kernel_full(height, width, kernel_size)
    if upsample:
        out_h = sh * (height - kh * sh * 2 - 1) + (1 + sh * 2) * kh
        out_w = sw * (width - kw * sw * 2 - 1) + (1 + sw * 2) * kw
        h_range, w_range = np.linspace(0, height, out_h), np.linspace(0, width, out_w)
    else:
        if kernel_size[0] == 1:
            h_range, w_range = np.arange(0, height, sh), np.arange(0, width, sw)
        else:
            h_range, w_range = _incre_range(height, kh, sh), _incre_range(width, kw, sw)
    lat_c = ((h_range / height) - 0.5) * np.pi
    lon_c = ((w_range / width) - 0.5) * (2 * np.pi)
    t = lat_c[:, None, None]
    lat = np.arcsin(cos_nu * np.sin(t) + ker_y * sin_nu * np.cos(t) / rho)
    lon = np.arctan(ker_x * sin_nu / (rho * np.cos(t) * cos_nu - ker_y * np.sin(t) * sin_nu))
    lon = lon[:, None, :, :] + lon_c[None, :, None, None]
    lat = (lat / np.pi + 0.5) * height
    lon = ((lon / (2 * np.pi) + 0.5) * width) % width
    H, W = len(lat_c), len(lon_c)
    out = np.empty((H, kh, W, kw, 2), dtype=np.float64)
    out[..., 0] = lat[:, :, None, :]
    out[..., 1] = lon.transpose(0, 2, 1, 3)
    return out.reshape(1, H * kh, W * kw, 2)


def full_sphere_grid(pattern, h, w):
    """SphereConv2d.genSamplingPattern (sphere_conv2d.py:35-47): (lat, lon) pixel positions -> (1, H*Kh, W*Kw, 2) float32
    grid in F.grid_sample convention (x = lon, y = lat)."""
    lat = (pattern[:, :, :, 0] / h) * 2 - 1
    lon = (pattern[:, :, :, 1] / w) * 2 - 1
    return np.stack((lon, lat), axis=-1).astype(np.float32)


_KEYS_ROW = ("p_x_st", "p_x_ed", "x_total", "y_total", "test_flag", "partial")
_KEYS_COL = ("p_y_st", "p_y_ed", "circular_flag")


class _FactorTable:
    """Growable device table of per-window factors: slot -> row of a (capacity, *shape) tensor."""

    def __init__(self, shape, dtype, device):
        self.shape, self.dtype, self.device = tuple(shape), dtype, device
        self.slots = {}
        self.data = torch.zeros((16,) + self.shape, dtype=dtype, device=device)

    def slot(self, key, build):
        s = self.slots.get(key)
        if s is not None:
            return s
        s = len(self.slots)
        if s >= self.data.shape[0]:
            grown = torch.zeros((2 * self.data.shape[0],) + self.shape, dtype=self.dtype, device=self.device)
            grown[:self.data.shape[0]] = self.data
            self.data = grown
        host = torch.from_numpy(np.ascontiguousarray(build())).to(self.dtype)
        if self.device.type == "cuda":
            host = host.pin_memory()  # pinned + non_blocking: the upload does not stall the host behind queued kernels
        self.data[s].copy_(host.view(self.shape), non_blocking=True)
        self.slots[key] = s
        return s


class WindowTables:
    """Factor tables of a FIXED family of training windows (every x start and every y start of the coordinate grid the
    sampler draws from), resident on the device: row slot = x-window index, column slot = y-window index.  With these,
    the grid of a batch is a pure device computation on two int32 index tensors (CUDA-graph capturable: no host work)."""

    def __init__(self, row_cps, col_cps, device):
        self.row_cps, self.col_cps, self.device = list(row_cps), list(col_cps), torch.device(device)
        self.y_total = float(self.row_cps[0]["y_total"])
        self._rows, self._cols = {}, {}

    def _row_tables(self, h):
        t = self._rows.get(h)
        if t is None:
            lat_n = np.empty((len(self.row_cps), h, 9), np.float32)
            lon = np.empty((len(self.row_cps), h, 9), np.float64)
            for i, cp in enumerate(self.row_cps):
                lat_g, lo = row_factor(h, cp)
                lat_n[i] = ((lat_g / cp["x_total"]) * 2 - 1).astype(np.float32).reshape(h, 9)
                lon[i] = lo.reshape(h, 9)
            t = self._rows[h] = (torch.from_numpy(lat_n).to(self.device), torch.from_numpy(lon).to(self.device))
        return t

    def _col_table(self, w):
        t = self._cols.get(w)
        if t is None:
            t = self._cols[w] = torch.from_numpy(np.stack([col_factor(w, cp) for cp in self.col_cps])).to(self.device)
        return t

    def prepare(self, sizes):
        for h in sizes:
            self._row_tables(h)
            self._col_table(h)

    def assemble(self, h, w, ix, iy):
        from . import lib
        import ctypes
        lat_n, lon = self._row_tables(h)
        nlon = self._col_table(w)
        B = ix.numel()
        out = torch.empty((B, 3 * h, 3 * w, 2), device=self.device, dtype=torch.float32)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(self.device):
            lib.call("spgan_sphere_grid_assemble", vp(out), vp(lat_n), vp(lon), vp(nlon), vp(ix), vp(iy), B, h, w,
                     self.y_total, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        return out


class DeviceWindows(list):
    """A training `coords_partial` (list of per-sample dicts, coord_handler.py:1027-1038) that also carries the windows'
    slot indices as device tensors: modules see the reference's list; GridCache builds the grids from `ix`, `iy` without
    touching the host.  The dicts describe the windows at construction time; after `ix` / `iy` are updated in place
    (CUDA-graph replay) only the device indices are authoritative."""

    def __init__(self, cps, tables, ix, iy):
        super().__init__(cps)
        self.tables, self.ix, self.iy = tables, ix, iy


class PositionGroup(list):
    """Test-mode `coords_partial` of SEVERAL lattice positions run as ONE batch (the reference's manager issues one
    generator call per position, test_managers/close_loop_infinite_generation.py:185-261): entry i is the dict of position
    i, whose `group` samples are rows [i*group, (i+1)*group) of the batch.  Every per-sample quantity of the generator is
    independent across the batch except the spherical conv's flat-concat channel table, which the reference builds per
    call — the table is therefore block-diagonal over the groups (functional._sphere_chan_map)."""

    def __init__(self, cps, group):
        super().__init__(cps)
        self.group = int(group)


class GridCache:
    """Device-resident sampling grids.

    Test mode (one coords_partial dict shared by the batch): whole (1, 3h, 3w, 2) grids keyed by (size, coords_partial),
    an LRU bounded in entries.  Training (a list of per-sample dicts): the grid separates into a row factor and a column
    factor (`row_factor`, `col_factor`); each distinct factor is computed once on the host, kept in a device table, and
    `spgan_sphere_grid_assemble` builds the (B, 3h, 3w, 2) batch grid from per-sample slot indices — the values are the
    host-built ones bit for bit, but a training step no longer pays ~0.3 ms of numpy and a blocking upload per sample
    and layer (the reference's models/spgan_ops_gs.py:767-781)."""

    def __init__(self, max_entries=8192):
        self.max_entries = max_entries
        self._store = OrderedDict()
        self._rows = {}
        self._cols = {}
        self.hits = 0
        self.misses = 0

    @staticmethod
    def _key(h, w, cp, device):
        return (h, w, str(device)) + tuple((k, cp.get(k, None)) for k in _KEYS)

    def get(self, h, w, cp, device):
        key = self._key(h, w, cp, device)
        g = self._store.get(key)
        if g is not None:
            self._store.move_to_end(key)
            self.hits += 1
            return g
        self.misses += 1
        g = torch.from_numpy(sampling_grid(h, w, cp)).to(device)
        self._store[key] = g
        if len(self._store) > self.max_entries:
            self._store.popitem(last=False)
        return g

    def assemble(self, h, w, cps, device):
        """(B, 3h, 3w, 2) grid of a list of per-sample coords_partial dicts, built on the device from factor tables."""
        from . import lib
        import ctypes
        device = torch.device(device)
        y_total = cps[0]["y_total"]
        if any(cp["y_total"] != y_total for cp in cps):
            return torch.cat([self.get(h, w, cp, device) for cp in cps], 0)
        rows = self._rows.get((h, str(device)))
        if rows is None:
            rows = self._rows[(h, str(device))] = (_FactorTable((h, 9), torch.float32, device),
                                                   _FactorTable((h, 9), torch.float64, device), {})
        cols = self._cols.get((w, str(device)))
        if cols is None:
            cols = self._cols[(w, str(device))] = _FactorTable((w,), torch.float64, device)
        lat_tab, lon_tab, memo = rows
        ix, iy = [], []
        for cp in cps:
            kr = tuple(cp.get(k, None) for k in _KEYS_ROW)
            if kr not in lat_tab.slots:
                lat_g, lon = row_factor(h, cp)
                memo[kr] = (((lat_g / cp["x_total"]) * 2 - 1).astype(np.float32).reshape(h, 9), lon.reshape(h, 9))
            s = lat_tab.slot(kr, lambda: memo[kr][0])
            s2 = lon_tab.slot(kr, lambda: memo[kr][1])
            assert s == s2
            memo.pop(kr, None)
            ix.append(s)
            kc = tuple(cp.get(k, None) for k in _KEYS_COL)
            iy.append(cols.slot(kc, lambda: col_factor(w, cp)))
        B = len(cps)
        idx = torch.tensor([ix, iy], dtype=torch.int32)
        if device.type == "cuda":
            idx = idx.pin_memory()
        idx = idx.to(device, non_blocking=True)
        out = torch.empty((B, 3 * h, 3 * w, 2), device=device, dtype=torch.float32)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(device):
            lib.call("spgan_sphere_grid_assemble", vp(out), vp(lat_tab.data), vp(lon_tab.data), vp(cols.data), vp(idx[0]),
                     vp(idx[1]), B, h, w, float(y_total), ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        return out

    def group_grid(self, h, w, cps, device):
        """(G, 3h, 3w, 2): one test-mode grid per lattice position of a PositionGroup (memoised per tuple of positions)."""
        key = ("group", h, w, str(device)) + tuple(self._key(h, w, cp, device) for cp in cps)
        g = self._store.get(key)
        if g is None:
            g = self._store[key] = torch.cat([self.get(h, w, cp, device) for cp in cps], 0).contiguous()
        return g

    def batch(self, h, w, coords_partial, batch, device):
        """Training: a list of per-sample dicts -> (B, 3h, 3w, 2); test: one dict -> (1, 3h, 3w, 2) shared by the
        batch (models/spgan_ops_gs.py:760-789)."""
        if isinstance(coords_partial, PositionGroup):
            if len(coords_partial) * coords_partial.group != batch:
                raise RuntimeError("PositionGroup of %d x %d samples for a batch of %d" % (len(coords_partial), coords_partial.group, batch))
            if len(coords_partial) == 1:
                return self.get(h, w, coords_partial[0], device)
            key = ("expanded", coords_partial.group, h, w, str(device)) + tuple(self._key(h, w, cp, device) for cp in coords_partial)
            g = self._store.get(key)
            if g is None:  # per-sample copy for the consumers that know only "one grid" or "one grid per sample"
                g = self._store[key] = self.group_grid(h, w, coords_partial, device).repeat_interleave(coords_partial.group, dim=0).contiguous()
            return g
        if isinstance(coords_partial, (list, tuple)):
            if len(coords_partial) != batch:
                raise RuntimeError("coords_partial has %d entries for a batch of %d" % (len(coords_partial), batch))
            if isinstance(coords_partial, DeviceWindows):
                return coords_partial.tables.assemble(h, w, coords_partial.ix, coords_partial.iy)
            if torch.device(device).type != "cuda":
                return torch.cat([self.get(h, w, cp, device) for cp in coords_partial], 0)
            return self.assemble(h, w, list(coords_partial), device)
        return self.get(h, w, coords_partial, device)


def assemble_reference(h, w, cp):
    """numpy statement of what spgan_sphere_grid_assemble computes from the factors (tests: equals sampling_grid)."""
    lat_g, lon = row_factor(h, cp)
    lat_n = ((lat_g / cp["x_total"]) * 2 - 1).astype(np.float32)
    y_total = float(cp["y_total"])
    a = lon[:, None, :, :] + col_factor(w, cp)[None, :, None, None]
    n = ((((a / 2.0) + 0.5) * y_total) / y_total) * 2.0 - 1.0
    out = np.empty((h, 3, w, 3, 2), dtype=np.float32)
    out[..., 0] = n.astype(np.float32).transpose(0, 2, 1, 3)
    out[..., 1] = lat_n[:, :, None, :]
    return out.reshape(1, 3 * h, 3 * w, 2)


GRID_CACHE = GridCache()
