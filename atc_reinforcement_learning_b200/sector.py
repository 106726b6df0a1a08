"""Sector compiler: Scenario -> flat arrays + derived constants + the exact MVA lookup grid, i.e. the host-side,
one-off part of the reference's AtcGym.__init__ (atc_gym.py:49-58, 88-110) and Corridor.__init__
(model.py:155-186).  Everything the kernels need is plain numpy here and is copied to the device by atc_create().

The MVA grid is an *exact* accelerator of Airspace.find_mva (model.py:282-292): a cell that no
polygon edge comes within `margin` of has one answer for all of its points, which is stored; any other cell stores
the set of polygons that could contain one of its points, and the kernel runs the reference's first-match
bbox + ray-cast scan over just that set.
"""
import math

import numpy as np

OBS_DIM = 10
MAX_MVA = 31
GRID_PAD = 2        # rings of cells around the bbox: the outer one is always "outside", so clamping is safe
LINE_EPS = 1e-9     # nm: closer than this to a cell's boundary line -> exact program (kernel uses the same value)


def ray_tracing_np(x, y, ring):
    """Vectorised restatement of model.py:318-337 for a closed ring (host side, used for the grid and faf_mva)."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    inside = np.zeros(x.shape, bool)
    n = len(ring)
    for i in range(1, n):
        p1x, p1y = ring[i - 1]
        p2x, p2y = ring[i]
        if p1y == p2y:
            continue        # y > min and y <= max cannot both hold
        cond = (y > min(p1y, p2y)) & (y <= max(p1y, p2y)) & (x <= max(p1x, p2x))
        xints = (y - p1y) * (p2x - p1x) / (p2y - p1y) + p1x
        inside ^= cond & ((p1x == p2x) | (x <= xints))
    return inside


def _rot_apply(phi_deg, vx, vy):
    """rot_matrix(phi) . [vx, vy]^T (model.py:345-348)"""
    phi = math.radians(phi_deg)
    c, s = math.cos(phi), math.sin(phi)
    return c * vx + s * vy, -s * vx + c * vy


class CompiledSector(object):
    def __init__(self, scenario, cell=0.25, margin=None, wind=None, grid_origin=None):
        mvas = scenario.mvas
        if len(mvas) > MAX_MVA:
            raise ValueError("at most %d MVA polygons are supported" % MAX_MVA)
        self.name = getattr(scenario, 'name', type(scenario).__name__)
        self.rings = [np.asarray(m.area_as_list, np.float64) for m in mvas]
        self.ring_xy = np.ascontiguousarray(np.concatenate(self.rings, 0))
        self.ring_off = np.cumsum([0] + [len(r) for r in self.rings]).astype(np.int32)
        self.mva_height = np.asarray([float(m.height) for m in mvas], np.float64)
        self.mva_bounds = np.ascontiguousarray(
            [[r[:, 0].min(), r[:, 1].min(), r[:, 0].max(), r[:, 1].max()] for r in self.rings], np.float64)
        b = self.mva_bounds
        self.bbox = np.asarray([b[:, 0].min(), b[:, 1].min(), b[:, 2].max(), b[:, 3].max()], np.float64)

        # ---- runway / corridor (model.py:155-186, 234-246)
        rw = scenario.runway
        self.runway = (float(rw.x), float(rw.y), float(rw.h))
        phi_from = rw.phi_from_runway
        self.phi_to_runway = float((phi_from + 180) % 360)
        self.normal = _rot_apply(phi_from, 0.0, 1.0)
        faf_dist, faf_angle, iaf_dist = 7.4, 45, 3
        corner_dist = iaf_dist / math.cos(math.radians(faf_angle))
        d = _rot_apply(phi_from, 0.0, faf_dist)
        self.faf = (rw.x + d[0], rw.y + d[1])
        inner = _rot_apply(phi_from, 0.0, corner_dist)
        c1 = _rot_apply(faf_angle, inner[0], inner[1])
        c2 = _rot_apply(-faf_angle, inner[0], inner[1])
        self.corner1 = (c1[0] + self.faf[0], c1[1] + self.faf[1])
        self.corner2 = (c2[0] + self.faf[0], c2[1] + self.faf[1])
        d = _rot_apply(phi_from, 0.0, faf_dist + iaf_dist)
        self.iaf = (rw.x + d[0], rw.y + d[1])
        self.tri_h = np.asarray([self.faf, self.corner1, self.corner2, self.faf], np.float64)
        self.tri_1 = np.asarray([self.faf, self.corner1, self.iaf, self.faf], np.float64)
        self.tri_2 = np.asarray([self.faf, self.corner2, self.iaf, self.faf], np.float64)
        self.sin_to_runway, self.cos_to_runway = _rot_apply(self.phi_to_runway, 0.0, 1.0)
        self.glide_tan = math.tan(3 * math.pi / 180)

        # ---- AtcGym.__init__ constants (atc_gym.py:49-58, 88-110)
        m = self.find_mva_np(np.asarray([self.faf[0]]), np.asarray([self.faf[1]]))[0]
        if m < 0:
            raise ValueError("the final approach fix lies outside the airspace")   # reference: ValueError too
        self.faf_mva = float(self.mva_height[m])
        lx, ly = self.bbox[2] - self.bbox[0], self.bbox[3] - self.bbox[1]
        self.world_max_distance = float(np.hypot(lx, ly))
        self.norm_min = np.asarray([self.bbox[0], self.bbox[1], 0, 0, 100, 0, 0, 0, -180, -180], np.float32)
        self.norm_max = np.asarray([lx, ly, 38000, 360, 200, 38000, 38000, self.world_max_distance, 360, 360],
                                   np.float32)

        # ---- entry points (scenarios.py:192-207)
        eps = scenario.entrypoints
        if len(eps) > 32:
            raise ValueError("at most 32 entry points are supported")
        self.entry_xyphi = np.ascontiguousarray([[e.x, e.y, e.phi] for e in eps], np.float64)
        self.level_off = np.cumsum([0] + [len(e.levels) for e in eps]).astype(np.int32)
        self.levels = np.concatenate([np.asarray(e.levels, np.int32) for e in eps]).astype(np.int32)
        for e in eps:
            for lv in e.levels:
                if not 0 <= lv * 100 <= 38000:
                    raise ValueError("invalid altitude")            # Airplane.__init__, model.py:35-36

        # ---- wind (extension)
        self.wind = None
        if wind is not None:
            w = np.ascontiguousarray(wind, np.float32)
            if w.ndim != 3 or w.shape[2] != 2 or w.shape[0] < 2 or w.shape[1] < 2:
                raise ValueError("wind must have shape [gy >= 2, gx >= 2, 2]")
            if not np.isfinite(w).all():
                raise ValueError("wind must be finite")
            self.wind = w

        self.cell = float(cell)
        # The kernel finds the grid cell of a point from its float32 coordinates (one FFMA per axis), so the cell it
        # picks may be the neighbour of the true one when the point is within the float32 error of a cell border.
        # Every cell's stored answer / program is therefore valid on the cell expanded by `margin`, chosen as 4x a
        # bound of that error: rounding of the coordinate, of the folded offset and of the FFMA result.
        if margin is None:
            n_cells = max(self.bbox[2] - self.bbox[0], self.bbox[3] - self.bbox[1]) / self.cell + 2 * GRID_PAD + 1
            err = 2.0 ** -24 * (float(np.abs(self.bbox).max()) + 2.0 * self.cell * n_cells)
            margin = max(1e-6, 4.0 * err)
        self.margin = float(margin)
        self.grid_origin = grid_origin
        self._build_grid()

    # ---------------------------------------------------------------------------------------------- reference scan
    def find_mva_np(self, x, y):
        """Airspace.find_mva (model.py:282-292) for arrays of points: first polygon index or -1."""
        x = np.asarray(x, np.float64)
        y = np.asarray(y, np.float64)
        out = np.full(x.shape, -1, np.int32)
        for m in range(len(self.rings) - 1, -1, -1):
            b = self.mva_bounds[m]
            hit = (b[0] <= x) & (x <= b[2]) & (b[1] <= y) & (y <= b[3])
            hit &= ray_tracing_np(x, y, self.rings[m])
            out[hit] = m
        return out

    # ---------------------------------------------------------------------------------------------- grid
    def _build_grid(self):
        cs, mg = self.cell, self.margin
        # the grid starts GRID_PAD cells outside the bbox and ends GRID_PAD cells beyond it
        x0, y0 = self.bbox[0] - GRID_PAD * cs, self.bbox[1] - GRID_PAD * cs
        if self.grid_origin is not None:                # a grid aligned with another one (sector.CompactGrid)
            if self.grid_origin[0] > x0 or self.grid_origin[1] > y0:
                raise ValueError("grid_origin must leave GRID_PAD cells outside the bbox")
            x0, y0 = float(self.grid_origin[0]), float(self.grid_origin[1])
        nx = int(math.floor((self.bbox[2] - x0) / cs)) + 1 + GRID_PAD
        ny = int(math.floor((self.bbox[3] - y0) / cs)) + 1 + GRID_PAD
        self.grid_nx, self.grid_ny = nx, ny
        self.grid_x0, self.grid_y0 = float(x0), float(y0)
        self.grid_inv_cell = 1.0 / cs
        edge_mask = np.zeros((ny, nx), np.uint32)
        touching = {}                 # (iy, ix) -> [(polygon, vertex index i)] of the edges that come near the cell
        for m, ring in enumerate(self.rings):
            for i in range(1, len(ring)):
                px, py = ring[i - 1]
                qx, qy = ring[i]
                # cells whose (margin-expanded) rectangle overlaps the segment's bbox
                ix0 = max(int(math.floor((min(px, qx) - mg - x0) / cs)) - 1, 0)
                ix1 = min(int(math.floor((max(px, qx) + mg - x0) / cs)) + 1, nx - 1)
                iy0 = max(int(math.floor((min(py, qy) - mg - y0) / cs)) - 1, 0)
                iy1 = min(int(math.floor((max(py, qy) + mg - y0) / cs)) + 1, ny - 1)
                if ix1 < ix0 or iy1 < iy0:
                    continue
                gx, gy = np.meshgrid(np.arange(ix0, ix1 + 1), np.arange(iy0, iy1 + 1))
                rx0, rx1 = x0 + gx * cs - mg, x0 + (gx + 1) * cs + mg
                ry0, ry1 = y0 + gy * cs - mg, y0 + (gy + 1) * cs + mg
                overlap = (rx0 <= max(px, qx)) & (rx1 >= min(px, qx)) & (ry0 <= max(py, qy)) & (ry1 >= min(py, qy))
                # separating axis = the segment's normal: all four corners strictly on one side -> no contact
                ex, ey = qx - px, qy - py
                tol = mg * (abs(ex) + abs(ey)) + 1e-12
                c = np.stack([ex * (cy - py) - ey * (cx - px) for cx, cy in
                              ((rx0, ry0), (rx1, ry0), (rx0, ry1), (rx1, ry1))])
                touch = overlap & ~((c.min(0) > tol) | (c.max(0) < -tol))
                edge_mask[iy0:iy1 + 1, ix0:ix1 + 1] |= np.where(touch, np.uint32(1 << m), np.uint32(0)).astype(np.uint32)
                for ty, tx in zip(*np.nonzero(touch)):
                    touching.setdefault((iy0 + int(ty), ix0 + int(tx)), []).append((m, i))
        # status of every polygon at the cell centres (valid for polygons with no edge near the cell)
        cxs = x0 + (np.arange(nx) + 0.5) * cs
        cys = y0 + (np.arange(ny) + 0.5) * cs
        gx, gy = np.meshgrid(cxs, cys)
        contain_mask = np.zeros((ny, nx), np.uint32)
        for m, ring in enumerate(self.rings):
            inside = ray_tracing_np(gx.ravel(), gy.ravel(), ring).reshape(ny, nx)
            contain_mask |= np.where(inside & ((edge_mask >> m) & 1 == 0), np.uint32(1 << m), np.uint32(0)).astype(np.uint32)
        # first polygon (list order) that contains the whole cell; later polygons can never be the first match
        lowest = contain_mask & (~contain_mask + np.uint32(1))           # lowest set bit (0 if none)
        upto = np.where(lowest > 0, (lowest << np.uint32(1)) - np.uint32(1), np.uint32(0x7FFFFFFF)).astype(np.uint32)
        cand = (edge_mask | contain_mask) & upto
        mixed = (edge_mask & upto) != 0
        first_idx = np.where(lowest > 0, np.log2(np.maximum(lowest, 1)).astype(np.int64) + 1, 0).astype(np.int64)
        self.grid_mixed_fraction = float(mixed.mean())

        # ---- per-cell programs for the mixed cells
        # For a candidate polygon only the edges that can interact with the cell's neighbourhood are kept:
        #   * an edge whose y-range misses the cell's (margin-expanded) strip, or that lies wholly left of the cell,
        #     can never satisfy the reference's crossing test (model.py:328-330) for a point of the cell -> dropped;
        #   * an edge wholly right of the cell whose y-range spans the whole strip always crosses -> folded into a
        #     parity bit;  everything else is tested exactly on the device.
        grid = np.zeros((ny, nx), np.uint16)
        uni = ~mixed
        grid[uni] = first_idx[uni].astype(np.uint16)
        prog, prog_off = [], []
        iys, ixs = np.nonzero(mixed)
        if len(iys) >= 0x8000:
            raise ValueError("too many mixed cells (%d): use a coarser grid_cell" % len(iys))
        edges = []      # per polygon: (global vertex index, ymin, ymax, xmin, xmax)
        for m, ring in enumerate(self.rings):
            el = []
            for i in range(1, len(ring)):
                (px, py), (qx, qy) = ring[i - 1], ring[i]
                if py == qy:
                    continue                                          # horizontal edges never cross
                el.append((int(self.ring_off[m]) + i, min(py, qy), max(py, qy), min(px, qx), max(px, qx)))
            edges.append(el)
        for k, (iy, ix) in enumerate(zip(iys.tolist(), ixs.tolist())):
            rx0, rx1 = x0 + ix * cs - mg, x0 + (ix + 1) * cs + mg
            ry0, ry1 = y0 + iy * cs - mg, y0 + (iy + 1) * cs + mg
            words, n_poly = [], 0
            cm, em = int(cand[iy, ix]), int(edge_mask[iy, ix])
            for m in range(len(self.rings)):
                if not (cm >> m) & 1:
                    continue
                n_poly += 1
                if not (em >> m) & 1:                                 # contains the whole cell: constant answer
                    words.append(m | (1 << 5))
                    continue
                parity, keep = 0, []
                for g, eymin, eymax, exmin, exmax in edges[m]:
                    if eymax < ry0 or eymin > ry1 or exmax < rx0:
                        continue
                    if exmin > rx1 and eymin < ry0 and eymax > ry1:
                        parity ^= 1
                        continue
                    keep.append(g)
                if len(keep) > 255:
                    raise ValueError("polygon with more than 255 edges near one cell")
                bb = self.mva_bounds[m]
                need_bbox = not (bb[0] < rx0 and bb[2] > rx1 and bb[1] < ry0 and bb[3] > ry1)
                words.append(m | (parity << 5) | (int(need_bbox) << 6) | (len(keep) << 8))
                words.extend(keep)
            off = len(prog)
            if off >= (1 << 26):
                raise ValueError("MVA grid programs too large")
            prog_off.append((n_poly << 26) | off)
            prog.extend(words)
            grid[iy, ix] = 0x8000 | k
        border = np.concatenate([grid[0], grid[-1], grid[:, 0], grid[:, -1]])
        if border.any():
            raise AssertionError("MVA grid: the outermost ring of cells must be uniformly outside")
        self.grid_cell = np.ascontiguousarray(grid)
        self.grid_prog_off = np.asarray(prog_off if prog_off else [0], np.uint32)
        self.grid_prog = np.asarray(prog if prog else [0], np.uint16)
        self.n_mixed = len(prog_off)

        # ---- single-line records: in most mixed cells every nearby edge lies on ONE line (a shared
        # polygon boundary crossing the cell, no vertex inside).  Polygon membership is then constant on each side of
        # that line, and for a point farther than LINE_EPS from it the reference's ray cast equals the true membership,
        # so the answer is a sign test; only points within LINE_EPS of the line run the exact program.
        rec = np.zeros((max(self.n_mixed, 1), 4), np.float64)
        rec_out = rec.view(np.int32).reshape(rec.shape[0], 8)      # out_pos, out_neg live in the 4th double's bits
        probe_k, probe_side, probe_xy = [], [], []
        n_line = 0
        for k, (iy, ix) in enumerate(zip(iys.tolist(), ixs.tolist())):
            tl = touching.get((iy, ix), [])
            if not tl:
                continue
            rx0, rx1 = x0 + ix * cs - mg, x0 + (ix + 1) * cs + mg
            ry0, ry1 = y0 + iy * cs - mg, y0 + (iy + 1) * cs + mg
            m0, i0 = tl[0]
            p, q = self.rings[m0][i0 - 1], self.rings[m0][i0]
            ln = math.hypot(q[0] - p[0], q[1] - p[1])
            if ln == 0.0:
                continue
            a, b = (q[1] - p[1]) / ln, -(q[0] - p[0]) / ln
            c = -(a * p[0] + b * p[1])
            ok = True
            for m, i in tl:
                for v in (self.rings[m][i - 1], self.rings[m][i]):
                    if abs(a * v[0] + b * v[1] + c) > 1e-9:                      # not on the line
                        ok = False
                    if rx0 - mg <= v[0] <= rx1 + mg and ry0 - mg <= v[1] <= ry1 + mg:   # a vertex inside the cell
                        ok = False
            if not ok:
                continue
            corners = [(rx0, ry0), (rx1, ry0), (rx0, ry1), (rx1, ry1)]
            d = [a * cx + b * cy + c for cx, cy in corners]
            for side, sel in ((6, max), (7, min)):
                dv = sel(d)
                if abs(dv) > 1e-7 and (dv > 0) == (side == 6):               # else: no point of the cell on that side
                    probe_k.append(k); probe_side.append(side); probe_xy.append(corners[d.index(dv)])
            rec[k, 0], rec[k, 1], rec[k, 2] = a, b, c
            n_line += 1
        if probe_xy:                                   # one vectorised reference scan for all probe corners
            pxy = np.asarray(probe_xy, np.float64)
            rec_out[np.asarray(probe_k), np.asarray(probe_side)] = self.find_mva_np(pxy[:, 0], pxy[:, 1]) + 1
        self.grid_line = np.ascontiguousarray(rec)
        self.line_fraction = n_line / max(self.n_mixed, 1)

    def cell_index_np(self, x, y):
        """The kernel's float32 cell index (mva_cell in csrc/atc_kernels.cu): fx = xf * scale + off clamped to the
        grid, truncated.  NaN clamps to cell 0 (outside).  numpy has no FMA: the product is formed in float64 and
        rounded once, which is what the FFMA does."""
        sc = np.float32(self.grid_inv_cell)
        ox = np.float32(-self.grid_x0 * self.grid_inv_cell)
        oy = np.float32(-self.grid_y0 * self.grid_inv_cell)
        with np.errstate(invalid='ignore'):
            xf = np.asarray(x, np.float64).astype(np.float32)
            yf = np.asarray(y, np.float64).astype(np.float32)
            fx = (xf.astype(np.float64) * np.float64(sc) + np.float64(ox)).astype(np.float32)
            fy = (yf.astype(np.float64) * np.float64(sc) + np.float64(oy)).astype(np.float32)
            fx = np.fmin(np.fmax(fx, np.float32(0)), np.float32(self.grid_nx - 1))
            fy = np.fmin(np.fmax(fy, np.float32(0)), np.float32(self.grid_ny - 1))
        return fx.astype(np.int64), fy.astype(np.int64)

    def lookup_np(self, x, y, cells=None):
        """Host restatement of the kernel's find_mva (grid + per-cell programs) — used by the CPU tests to check the
        accelerator against the brute-force reference scan.  `cells` = (ix, iy) overrides the cell choice (the tests
        use it to show that every cell within `margin` of a point gives the same, exact answer)."""
        x = np.asarray(x, np.float64)
        y = np.asarray(y, np.float64)
        out = np.full(x.shape, -1, np.int32)
        if cells is None:
            ix, iy = self.cell_index_np(x, y)
        else:
            ix, iy = cells
        cell = self.grid_cell[iy, ix]
        uniform = (cell & 0x8000) == 0
        out[uniform] = cell[uniform].astype(np.int32) - 1
        ring = self.ring_xy
        rec_out = self.grid_line.view(np.int32).reshape(-1, 8)
        for i in np.nonzero(~uniform)[0]:
            k = int(cell[i]) & 0x7FFF
            a, b, c = self.grid_line[k, :3]
            if a != 0.0 or b != 0.0:
                d = a * float(x[i]) + b * float(y[i]) + c
                if d > LINE_EPS:
                    out[i] = rec_out[k, 6] - 1
                    continue
                if d < -LINE_EPS:
                    out[i] = rec_out[k, 7] - 1
                    continue
            po = int(self.grid_prog_off[k])
            n_poly, p = po >> 26, po & 0x3FFFFFF
            px, py = float(x[i]), float(y[i])
            for _ in range(n_poly):
                h = int(self.grid_prog[p]); p += 1
                m, par, need_bbox, ne = h & 31, (h >> 5) & 1, (h >> 6) & 1, h >> 8
                ok = True
                if need_bbox:
                    b = self.mva_bounds[m]
                    ok = b[0] <= px <= b[2] and b[1] <= py <= b[3]
                for j in range(ne):
                    g = int(self.grid_prog[p + j])
                    p1x, p1y = ring[g - 1]
                    p2x, p2y = ring[g]
                    if py > min(p1y, p2y) and py <= max(p1y, p2y) and px <= max(p1x, p2x):
                        xints = (py - p1y) * (p2x - p1x) / (p2y - p1y) + p1x
                        if p1x == p2x or px <= xints:
                            par ^= 1
                p += ne
                if ok and par:
                    out[i] = m
                    break
        return out


# ------------------------------------------------------------------------------------------------ compact (shared-memory) grid
COMPACT_ESCAPE = 127          # line id meaning "not decidable here": take the fine grid
COMPACT_SUB0 = 126            # coarse level only: line ids 126 / 127 = "look in sub-block (id & 1) * 256 + bits 7-14"
COMPACT_MAX_LINES = 126
COMPACT_MAX_ANSWER = 15       # polygon index + 1 must fit four bits
COMPACT_SUB = 8               # a sub-block refines one coarse cell into 8 x 8 cells
COMPACT_MAX_BLOCKS = 511      # block 511 is the shared "undecidable everywhere" block


def _compact_cells(cs, lines, index):
    """CompiledSector grid -> compact u16 cells (uniform / single-line / 0x8000 | 127), line table shared via `index`."""
    g = cs.grid_cell
    out = np.where((g & 0x8000) == 0, g, np.uint16(0x8000 | COMPACT_ESCAPE)).astype(np.uint16)
    rec = cs.grid_line
    rec_out = rec.view(np.int32).reshape(-1, 8)
    iys, ixs = np.nonzero((g & 0x8000) != 0)
    for iy, ix in zip(iys.tolist(), ixs.tolist()):
        k = int(g[iy, ix]) & 0x7FFF
        a, b, c = (float(v) for v in rec[k, :3])
        if a == 0.0 and b == 0.0:
            continue
        pos, neg = int(rec_out[k, 6]), int(rec_out[k, 7])
        if not (0 <= pos <= COMPACT_MAX_ANSWER and 0 <= neg <= COMPACT_MAX_ANSWER):
            continue
        key = (a, b, c)
        lid = index.get(key)
        if lid is None:
            if len(lines) >= COMPACT_MAX_LINES:
                continue                                          # table full: the cell stays undecidable
            lid = index[key] = len(lines)
            lines.append((a, b, c, 0.0))
        out[iy, ix] = 0x8000 | lid | (pos << 7) | (neg << 11)
    return out


class CompactGrid(object):
    """A second, coarse MVA grid small enough for the shared memory of one SM: the rollout kernel
    with one CTA per SM keeps it next to its message rings, so the per-step lookup is a shared-memory load instead of
    an L2 round trip.  Built from the same exact machinery as the fine grid (CompiledSector at the coarse cell and at
    1/8 of it, same origin), two levels:

      coarse cell (u16)  bit 15 clear: polygon index + 1 of the whole (margin-grown) cell, 0 = outside
                         bit 15 set, line id (bits 0-6) < 126: ONE boundary line crosses the cell; bits 7-10 / 11-14 =
                                     answer (polygon index + 1) on its positive / negative side
                         bit 15 set, line id 126 / 127: the cell holds a vertex or several lines; its 8 x 8 sub-block
                                     number (id & 1) * 256 + bits 7-14 refines it (block 511: nothing decidable)
      sub-block cells    same encoding at 1/8 of the cell size; line id 127 = undecidable
      lines (f64)        [n_lines][4]: a, b, c (a*a + b*b = 1), 0 — deduplicated over both levels

    The index is taken at the sub-cell resolution (float32, like the fine grid's) and shifted down by 3 for the coarse
    cell.  A point farther than LINE_EPS from its cell's line takes that side's answer; everything undecidable is
    resolved by the fine grid, which is exact everywhere.  Only sectors with at most 126 distinct boundary lines and 15
    polygons get a compact grid (LOWW: 12 polygons, 60 lines)."""

    def __init__(self, scenario, cell, wind=None):
        sub = COMPACT_SUB
        probe = CompiledSector(scenario, cell=cell)                      # fixes the common origin
        origin = (probe.grid_x0, probe.grid_y0)
        fine = CompiledSector(scenario, cell=cell / sub, grid_origin=origin)
        cs = CompiledSector(scenario, cell=cell, margin=max(probe.margin, fine.margin))
        assert (cs.grid_x0, cs.grid_y0) == origin
        self.cell, self.margin = cs.cell, cs.margin
        self.grid_nx, self.grid_ny = cs.grid_nx, cs.grid_ny
        self.grid_x0, self.grid_y0, self.grid_inv_cell = cs.grid_x0, cs.grid_y0, cs.grid_inv_cell
        self._fine = fine                                                 # cell_index_np at the sub-cell resolution
        self.sub_inv_cell = fine.grid_inv_cell
        lines, index = [], {}
        coarse = _compact_cells(cs, lines, index)
        fcells = _compact_cells(fine, lines, index)
        ny, nx = coarse.shape
        pad = np.zeros((ny * sub, nx * sub), np.uint16)                   # beyond the fine grid: outside
        fy, fx = min(pad.shape[0], fcells.shape[0]), min(pad.shape[1], fcells.shape[1])
        pad[:fy, :fx] = fcells[:fy, :fx]
        blocks = []
        for iy, ix in zip(*[v.tolist() for v in np.nonzero(coarse == (0x8000 | COMPACT_ESCAPE))]):
            j = len(blocks)
            if j >= COMPACT_MAX_BLOCKS:
                j = COMPACT_MAX_BLOCKS
            else:
                blocks.append(pad[iy * sub:(iy + 1) * sub, ix * sub:(ix + 1) * sub].reshape(-1))
            coarse[iy, ix] = 0x8000 | (COMPACT_SUB0 + (j >> 8)) | ((j & 255) << 7)
        self.n_blocks = len(blocks)
        self.n_coarse = nx * ny
        flat = [coarse.reshape(-1)] + blocks
        self.grid_cell = np.ascontiguousarray(np.concatenate(flat).astype(np.uint16))
        self.coarse = coarse
        self.lines = np.ascontiguousarray(lines if lines else [(0.0, 0.0, 0.0, 0.0)], np.float64)
        self.n_lines = len(lines)
        self.mixed_fraction = float(((coarse & 0x8000) != 0).mean())
        self.block_fraction = float((((coarse & 0x8000) != 0) & ((coarse & 127) >= COMPACT_SUB0)).mean())
        sub_esc = sum(int((b == (0x8000 | COMPACT_ESCAPE)).sum()) for b in blocks)
        self.escape_fraction = sub_esc / float(sub * sub * max(self.n_coarse, 1))      # area that needs the fine grid
        self.nbytes = int(self.grid_cell.nbytes + self.lines.nbytes)

    def cell_index_np(self, x, y):
        """(ix8, iy8): the kernel's float32 index at the sub-cell resolution, clamped to the compact grid."""
        ix8, iy8 = self._fine.cell_index_np(x, y)
        return (np.minimum(ix8, self.grid_nx * COMPACT_SUB - 1), np.minimum(iy8, self.grid_ny * COMPACT_SUB - 1))

    def lookup_np(self, x, y, fine, cells=None):
        """Host restatement of the kernel's shared-memory lookup; `fine` is the CompiledSector whose exact lookup
        resolves the undecidable points; `cells` = (ix8, iy8) overrides the index.  Returns (polygon index or -1, mask
        of the points the fine grid resolved)."""
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        ix8, iy8 = self.cell_index_np(x, y) if cells is None else cells
        g = self.grid_cell.astype(np.int64)
        cell = g[(iy8 >> 3) * self.grid_nx + (ix8 >> 3)]
        is_blk = ((cell & 0x8000) != 0) & ((cell & 127) >= COMPACT_SUB0)
        j = ((cell & 1) << 8) | ((cell >> 7) & 255)
        j = np.where(is_blk & (j < self.n_blocks), j, 0)
        sub_cell = g[np.minimum(self.n_coarse + 64 * j + ((iy8 & 7) << 3) + (ix8 & 7), len(g) - 1)] if self.n_blocks else cell
        undecidable_blk = is_blk & ((((cell & 1) << 8) | ((cell >> 7) & 255)) >= self.n_blocks)
        cell = np.where(is_blk, sub_cell, cell)
        out = np.where((cell & 0x8000) == 0, cell - 1, -2)
        lid = cell & 127
        has_line = ((cell & 0x8000) != 0) & (lid < COMPACT_SUB0) & ~undecidable_blk
        out = np.where(undecidable_blk, -2, out)
        ln = self.lines[np.where(has_line, lid, 0)]
        # the kernel evaluates fma(a, x, fma(b, y, c)); numpy has no FMA, but d is only compared against
        # LINE_EPS = 1e-9, far above the rounding of either order
        d = ln[:, 0] * x + (ln[:, 1] * y + ln[:, 2])
        out = np.where(has_line & (d > LINE_EPS), ((cell >> 7) & 15) - 1, out)
        out = np.where(has_line & (d < -LINE_EPS), ((cell >> 11) & 15) - 1, out)
        slow = out == -2
        if slow.any():
            out = out.copy()
            out[slow] = fine.lookup_np(x[slow], y[slow])
        return out.astype(np.int32), slow


def build_compact_grid(scenario, budget_bytes, cells=(0.25, 0.3, 0.35, 0.4, 0.5, 0.6, 0.75, 1.0)):
    """The finest CompactGrid of `cells` that fits `budget_bytes`, or None (too many polygons, or nothing fits)."""
    if len(scenario.mvas) > COMPACT_MAX_ANSWER:
        return None
    rings = [np.asarray(m.area_as_list, np.float64) for m in scenario.mvas]
    xs = np.concatenate([r[:, 0] for r in rings]); ys = np.concatenate([r[:, 1] for r in rings])
    for c in cells:
        nx = int(math.floor((xs.max() - xs.min()) / c)) + 2 + 2 * GRID_PAD      # a cheap size estimate before the build
        ny = int(math.floor((ys.max() - ys.min()) / c)) + 2 + 2 * GRID_PAD
        if nx * ny * 2 > budget_bytes:
            continue
        try:
            g = CompactGrid(scenario, c)
        except ValueError:
            continue
        if g.nbytes <= budget_bytes:
            return g
    return None
