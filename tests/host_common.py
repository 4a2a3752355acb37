"""Shared helpers for the host-mirror tests (CPU): feed the same mesh to the oracle and to
the C++ host mirror (xyst_b200/host) and compare what they build."""
import numpy as np
import oraclelib as O
from xyst_b200 import hostapi as H


def fixture_to_host_mesh(mesh):
    """Side sets of a regression fixture as boundary-triangle lists (global ids)."""
    bt = mesh["block_type"]; bn = mesh["block_n"]
    tris = mesh["tris"]; tets = mesh["tets"]
    expofa = np.array([[0, 1, 3], [1, 2, 3], [0, 3, 2], [0, 2, 1]])
    # file-internal element id -> (type, block-relative id), ExodusIIMeshReader.cpp:697-735
    def blkrel(i):
        e = 0; ntri = 0; ntet = 0
        for t, n in zip(bt, bn):
            e += int(n)
            if e > i:
                return (0, i - ntet) if t == 0 else (1, i - ntri)
            if t == 0:
                ntri += int(n)
            else:
                ntet += int(n)
        raise IndexError(i)
    off = [0]; out = []
    for s in range(len(mesh["set_id"])):
        a, b = int(mesh["set_off"][s]), int(mesh["set_off"][s + 1])
        for i in range(a, b):
            ty, r = blkrel(int(mesh["set_elem"][i]))
            if ty == 0:
                out.append(tris[r])
            else:
                out.append(tets[r][expofa[int(mesh["set_side"][i])]])
        off.append(len(out))
    return dict(coord=mesh["coord"], tets=tets, set_id=mesh["set_id"],
                set_off=np.asarray(off, np.uint64), set_tri=np.asarray(out, np.uint64).reshape(-1, 3))


def host_mesh_to_oracle(m):
    """A host mesh (tets + side-set triangles) as oracle input: the triangles become a TRI
    element block and the side sets refer to it, as in the reference's regression meshes."""
    ntri = len(m["set_tri"])
    return dict(coord=m["coord"], tets=m["tets"], tris=m["set_tri"],
                block_type=np.array([0, 1], np.int32), block_n=np.array([ntri, len(m["tets"])], np.uint64),
                set_id=m["set_id"], set_off=m["set_off"], set_elem=np.arange(ntri, dtype=np.uint64),
                set_side=np.zeros(ntri, np.uint64))


def edge_dict(get):
    """{(p,q) -> (dx,dy,dz)} over all superedge groups, orientation p->q as stored."""
    lpoed = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]; lpoet = [(0, 1), (1, 2), (2, 0)]
    d = {}
    for k, (nn, lp) in enumerate([(4, lpoed), (3, lpoet), (2, [(0, 1)])]):
        se = get("dsupedge%d" % k).reshape(-1, nn).astype(np.int64)
        si = get("dsupint%d" % k).reshape(len(se), len(lp), 3)
        for e in range(len(se)):
            for j, (a, b) in enumerate(lp):
                p, q = int(se[e, a]), int(se[e, b])
                v = si[e, j]
                if p > q:
                    p, q, v = q, p, -v
                assert (p, q) not in d
                d[(p, q)] = tuple(v)
    return d


def face_multiset(triinpoel, bface_flat):
    """{set id -> sorted list of (rotation-normalised) oriented faces}."""
    tri = np.asarray(triinpoel).reshape(-1, 3).astype(np.int64)
    out = {}; i = 0; f = np.asarray(bface_flat).astype(np.int64)
    while i < len(f):
        s, n = int(f[i]), int(f[i + 1]); ids = f[i + 2:i + 2 + n]; i += 2 + n
        faces = []
        for t in tri[ids]:
            k = int(np.argmin(t)); faces.append(tuple(int(x) for x in np.roll(t, -k)))
        out[s] = sorted(faces)
    return out
