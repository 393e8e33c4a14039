"""Cloth template: the triangle mesh behind `p.loadCloth` / `p.clothParams` (reference envs/dressing.py:146-154).

Host-side preparation of the immutable cloth model that `ag_cloth_init` uploads: nodes, the link (edge) list in
*colour order*, node -> face adjacency for the node normals, node areas.  The solver semantics restate Bullet's
`btSoftBody` position solver (recalled); this module only prepares the topology.

Link order.  Bullet relaxes links one after the other in list order (Gauss-Seidel).  Here the list is sorted by an edge
colouring (links of one colour share no node), so a sequential sweep over the list -- what the CPU oracle does -- and a
colour-by-colour parallel sweep -- what the CUDA kernel does -- are the same computation.
"""
import os

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')

# p.clothParams defaults of the reference call (dressing.py:147); Bullet's own defaults differ (kDF 0.2, kKHR 0.1, kAHR 0.7)
DRESSING_PARAMS = dict(kLST=0.055, kDP=0.01, kDG=10.0, kLF=0.0, kDF=0.39, kCHR=1.0, kKHR=1.0, kAHR=1.0, piterations=5,
                       margin=0.04, air_density=1.2, total_mass=0.16)


def _edge_colouring(links, n_nodes, cap=1024):
    """Balanced greedy edge colouring in list order: among the colours not used at either end node, the one holding the fewest links.
    The number of colours K starts at max(maximum node degree, ceil(links / cap)) and grows until every colour holds <= `cap` links:
    the CUDA kernel relaxes a colour with one link per thread of a 1024-thread CTA, so the gown's 11 640 links become 12 colours of
    970 = 12 passes of the block (a first-fit colouring gives 10 colours, six of them with ~1 900 links = 16 passes).  Deterministic."""
    deg = np.bincount(np.asarray(links).ravel(), minlength=n_nodes)
    K = max(int(deg.max()), -(-len(links) // cap), 1)
    while True:
        used = np.zeros((n_nodes, K), dtype=bool)
        count = np.zeros(K, dtype=np.int64)
        col = np.empty(len(links), dtype=np.int32)
        ok = True
        for i, (a, b) in enumerate(links):
            free = ~(used[a] | used[b]) & (count < cap)
            if not free.any():
                ok = False
                break
            c = int(np.argmin(np.where(free, count, 1 << 60)))
            col[i] = c
            used[a, c] = used[b, c] = True
            count[c] += 1
        if ok:
            return col
        K += 1


def _bank_friendly(L):
    """Order (and orient) the independent links of one colour so that every aligned group of 8 consecutive links touches 8 different
    16-byte bank groups of shared memory on the i side and on the j side (node index mod 8): a quarter-warp then reads / writes its 8
    float4 positions in one wavefront.  Greedy first fit; relaxing a link is symmetric in its two nodes, so orientation is free."""
    rem = [tuple(int(v) for v in x) for x in L]
    out = []
    while rem:
        ua = ub = 0
        group, keep = [], []
        for (i, j) in rem:
            if len(group) < 8:
                a, b = i & 7, j & 7
                if not (ua >> a) & 1 and not (ub >> b) & 1:
                    group.append((i, j)); ua |= 1 << a; ub |= 1 << b
                    continue
                if not (ua >> b) & 1 and not (ub >> a) & 1:
                    group.append((j, i)); ua |= 1 << b; ub |= 1 << a
                    continue
            keep.append((i, j))
        while len(group) < 8 and keep:
            group.append(keep.pop(0))
        out += group
        rem = keep
    return np.array(out, dtype=np.int32).reshape(-1, 2)


def _locality_order(x, links, n_nodes):
    """Breadth-first node order over the mesh graph from the node with the smallest x coordinate (ties: index): nodes that
    share a link get nearby internal indices, which keeps a warp's shared-memory accesses of one colour clustered."""
    adj = [[] for _ in range(n_nodes)]
    for a, b in links:
        adj[a].append(b)
        adj[b].append(a)
    seen = np.zeros(n_nodes, dtype=bool)
    order = []
    for start in np.lexsort((np.arange(n_nodes), x[:, 0])):
        if seen[start]:
            continue
        seen[start] = True
        queue = [int(start)]
        while queue:
            nxt = []
            for u in queue:
                order.append(u)
                for w in sorted(adj[u]):
                    if not seen[w]:
                        seen[w] = True
                        nxt.append(w)
            queue = nxt
    return np.array(order, dtype=np.int32)


class ClothModel:
    """Immutable cloth template.  Public node ids are Bullet's (see tools/compile_assets.compile_cloth); the device works
    on an internal permutation (`order[internal] = public`, `rank[public] = internal`)."""

    def __init__(self, verts, faces, scale=1.0, params=None, reorder=True):
        self.params = dict(DRESSING_PARAMS)
        if params:
            self.params.update(params)
        self.n_nodes = int(len(verts))
        self.rest = np.asarray(verts, dtype=np.float64) * float(scale)      # public order, local frame
        faces = np.asarray(faces, dtype=np.int32)
        e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
        e = np.unique(np.sort(e, axis=1), axis=0)                           # unique undirected edges, lexicographic
        self.order = _locality_order(self.rest, e, self.n_nodes) if reorder else np.arange(self.n_nodes, dtype=np.int32)
        self.rank = np.empty(self.n_nodes, dtype=np.int32)
        self.rank[self.order] = np.arange(self.n_nodes, dtype=np.int32)
        # ---- everything below is in INTERNAL node ids
        self.faces = self.rank[faces]
        links = np.sort(self.rank[e], axis=1)
        links = links[np.lexsort((links[:, 1], links[:, 0]))]
        col = _edge_colouring(links, self.n_nodes)
        idx = np.lexsort((links[:, 1], links[:, 0], col))                   # colour-major, then by nodes
        links = links[idx]
        self.link_colour = col[idx]
        self.n_colours = int(col.max()) + 1
        self.colour_off = np.searchsorted(self.link_colour, np.arange(self.n_colours + 1)).astype(np.int32)
        # within a colour the order is free (the links share no node): arrange it for conflict-free shared-memory access
        self.links = np.ascontiguousarray(np.concatenate([_bank_friendly(links[self.colour_off[c]:self.colour_off[c + 1]]) for c in range(self.n_colours)]))
        xr = self.rest[self.order]
        d = xr[self.links[:, 1]] - xr[self.links[:, 0]]
        self.link_rest2 = np.einsum('ij,ij->i', d, d)                       # Bullet m_c1 = rest length squared
        # node areas (Bullet updateArea: a third of the adjacent face areas, from the initial configuration)
        a, b, c = xr[self.faces[:, 0]], xr[self.faces[:, 1]], xr[self.faces[:, 2]]
        farea = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        self.node_area = np.zeros(self.n_nodes)
        for k in range(3):
            np.add.at(self.node_area, self.faces[:, k], farea / 3.0)
        # node -> adjacent faces as (next, next-next) pairs in the face's winding, ascending face id
        lst = [[] for _ in range(self.n_nodes)]
        for f in self.faces:
            for k in range(3):
                lst[f[k]].append((f[(k + 1) % 3], f[(k + 2) % 3]))
        self.nf_off = np.zeros(self.n_nodes + 1, dtype=np.int32)
        self.nf_off[1:] = np.cumsum([len(l) for l in lst])
        self.nf_pair = np.array([p for l in lst for p in l], dtype=np.int32).reshape(-1, 2)
        self.inv_mass = self.n_nodes / float(self.params['total_mass'])     # uniform (Bullet setTotalMass, not from faces)

    @staticmethod
    def load(name='hospitalgown_reduced', scale=1.4, params=None, reorder=True):
        z = np.load(os.path.join(ASSET_DIR, name + '.agcloth.npz'))
        return ClothModel(z['verts'], z['faces'], scale=scale, params=params, reorder=reorder)

    # ---- placement (reference dressing.py:146: scale, position, orientation; the position is scaled too)
    def place(self, position, quat_xyzw=(0, 0, 0, 1), scale_position=1.0):
        """World positions [n_nodes][3] in PUBLIC order for x = R * rest + scale_position * position."""
        x, y, z, w = quat_xyzw
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        return self.rest @ R.T + np.asarray(position, dtype=np.float64) * scale_position

    def to_internal(self, a_public):
        """[..., n_nodes, C] public order -> internal order."""
        return np.take(a_public, self.order, axis=-2)

    def to_public(self, a_internal):
        return np.take(a_internal, self.rank, axis=-2)
