"""Random sector generator for the tests: a jittered quad mesh over a rectangle (shared vertices, so neighbouring MVA
polygons share their boundary edges exactly like the reference's LOWW data), random heights, a runway whose final
approach fix lies inside the airspace, a few entry points.  Returns a sector document in the 'atc-b200-sector/1' format."""
import numpy as np


def random_sector(seed, nx=4, ny=3, size=(60.0, 45.0), holes=True):
    rng = np.random.RandomState(seed)
    xs = np.linspace(0.0, size[0], nx + 1)
    ys = np.linspace(0.0, size[1], ny + 1)
    vx, vy = np.meshgrid(xs, ys)
    jit = 0.3 * min(size[0] / nx, size[1] / ny)
    vx = vx + rng.uniform(-jit, jit, vx.shape) * (np.arange(nx + 1)[None, :] % nx != 0)      # keep the outline straight
    vy = vy + rng.uniform(-jit, jit, vy.shape) * (np.arange(ny + 1)[:, None] % ny != 0)
    vx, vy = np.round(vx, 2), np.round(vy, 2)                  # 2-decimal coordinates like the reference data
    mvas = []
    for j in range(ny):
        for i in range(nx):
            if holes and rng.uniform() < 0.12:
                continue                                        # a hole: "outside" in the middle of the bbox
            quad = [(vx[j, i], vy[j, i]), (vx[j, i + 1], vy[j, i + 1]), (vx[j + 1, i + 1], vy[j + 1, i + 1]),
                    (vx[j + 1, i], vy[j + 1, i])]
            if rng.uniform() < 0.5:
                quad = quad[::-1]                               # mixed orientations
            if rng.uniform() < 0.4:                             # extra vertex on an edge (T-junction free: own edge only)
                k = rng.randint(4)
                p, q = quad[k], quad[(k + 1) % 4]
                # only on the outer outline, where no neighbour shares the edge
                if (p[0] == q[0] and p[0] in (0.0, size[0])) or (p[1] == q[1] and p[1] in (0.0, size[1])):
                    quad.insert(k + 1, (round(0.5 * (p[0] + q[0]), 2), round(0.5 * (p[1] + q[1]), 2)))
            ring = [[float(x), float(y)] for x, y in quad]
            ring.append(ring[0])
            mvas.append({'height': int(rng.randint(20, 80)) * 100, 'ring': ring})
    rng.shuffle(mvas)                                           # list order matters (first match wins)
    doc = {'format': 'atc-b200-sector/1', 'name': 'random-%d' % seed, 'mvas': mvas,
           'runway': {'x': float(size[0] * rng.uniform(0.35, 0.65)), 'y': float(size[1] * rng.uniform(0.35, 0.65)),
                      'h': float(rng.randint(0, 10) * 100), 'phi_from_runway': float(rng.randint(0, 36) * 10)}}
    eps = []
    for _ in range(9):
        eps.append({'x': float(np.round(rng.uniform(2, size[0] - 2), 1)), 'y': float(np.round(rng.uniform(2, size[1] - 2), 1)),
                    'phi': float(rng.randint(0, 36) * 10), 'levels': [int(v) for v in rng.choice(np.arange(100, 260, 10), 4)]})
    doc['entrypoints'] = eps[:1]
    doc['entrypoints_random'] = eps
    return doc
