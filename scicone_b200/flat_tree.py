"""Flat (C-ABI) encoding of a SCICoNE tree: ``scicone_tree`` of include/scicone_score.h.

Mirrors the fields of the reference's ``Node`` that feed the attachment score
(reference: scicone/src/Node.h:18-81 -- ``id``, parent link, ``c_change``, ``c``)
and ``Tree::nu`` (scicone/src/Tree.h:52).  Nodes are stored parent-before-child,
entry 0 is the root.
"""
import ctypes as C

import numpy as np


class CTree(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("node_id", C.POINTER(C.c_int32)),
        ("parent", C.POINTER(C.c_int32)),
        ("chg_off", C.POINTER(C.c_int32)),
        ("chg_region", C.POINTER(C.c_int32)),
        ("chg_value", C.POINTER(C.c_int32)),
        ("gen_off", C.POINTER(C.c_int32)),
        ("gen_region", C.POINTER(C.c_int32)),
        ("gen_value", C.POINTER(C.c_int32)),
        ("nu", C.c_double),
    ]


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class FlatTree:
    """Host-side tree: ``nodes`` is a list of dicts ``{"id", "parent" (id or -1), "c_change": {region: value}}``.

    Genotypes ``c`` are derived as the reference's ``Tree::update_label`` does
    (scicone/src/Tree.h:460-475): parent genotype plus events, zeros erased --
    unless the caller supplies ``"c"`` explicitly.
    """

    def __init__(self, nodes, nu):
        order = self._topological(nodes)
        self.nu = float(nu)
        self.ids = [n["id"] for n in order]
        index = {nid: i for i, nid in enumerate(self.ids)}
        if len(index) != len(order):
            raise ValueError("duplicate node ids")
        self.parent = [index[n["parent"]] if n["parent"] is not None and n["parent"] >= 0 else -1 for n in order]
        self.c_change = [dict(sorted((int(k), int(v)) for k, v in dict(n["c_change"]).items())) for n in order]
        self.c = []
        for i, n in enumerate(order):
            if "c" in n and n["c"] is not None:
                g = dict(sorted((int(k), int(v)) for k, v in dict(n["c"]).items()))
            elif self.parent[i] < 0:
                g = {}
            else:
                g = dict(self.c[self.parent[i]])
                for k, v in self.c_change[i].items():
                    nv = g.get(k, 0) + v
                    if nv == 0:
                        g.pop(k, None)
                    else:
                        g[k] = nv
                g = dict(sorted(g.items()))
            self.c.append(g)
        self._pack()

    @staticmethod
    def _topological(nodes):
        by_id = {n["id"]: n for n in nodes}
        children = {}
        root = None
        for n in nodes:
            p = n["parent"]
            if p is None or p < 0:
                root = n
            else:
                children.setdefault(p, []).append(n)
        if root is None:
            raise ValueError("tree has no root")
        out, stack = [], [root]
        while stack:
            top = stack.pop()
            out.append(top)
            stack.extend(children.get(top["id"], []))
        if len(out) != len(by_id):
            raise ValueError("tree is not connected")
        return out

    def _pack(self):
        n = len(self.ids)
        self._node_id = _i32(self.ids)
        self._parent = _i32(self.parent)
        chg_off, chg_r, chg_v, gen_off, gen_r, gen_v = [0], [], [], [0], [], []
        for i in range(n):
            for k, v in self.c_change[i].items():
                chg_r.append(k)
                chg_v.append(v)
            chg_off.append(len(chg_r))
            for k, v in self.c[i].items():
                gen_r.append(k)
                gen_v.append(v)
            gen_off.append(len(gen_r))
        self._chg_off, self._chg_r, self._chg_v = _i32(chg_off), _i32(chg_r + [0]), _i32(chg_v + [0])
        self._gen_off, self._gen_r, self._gen_v = _i32(gen_off), _i32(gen_r + [0]), _i32(gen_v + [0])
        self.ctree = CTree(n, _ptr(self._node_id), _ptr(self._parent), _ptr(self._chg_off), _ptr(self._chg_r),
                           _ptr(self._chg_v), _ptr(self._gen_off), _ptr(self._gen_r), _ptr(self._gen_v), self.nu)

    @property
    def n_nodes(self):
        return len(self.ids)

    def index_of(self, node_id):
        return self.ids.index(node_id)

    def subtree_ids(self, node_id):
        idx = self.index_of(node_id)
        inside = {idx}
        for i in range(idx + 1, self.n_nodes):
            if self.parent[i] in inside:
                inside.add(i)
        return sorted(self.ids[i] for i in inside)

    def byref(self):
        return C.byref(self.ctree)

    @classmethod
    def from_json(cls, tree_json):
        """From the ``{"nu":..,"nodes":[{"id","parent","c_change":[[k,v]..],"c":[[k,v]..]}]}`` dumps under tests/golden."""
        nodes = [{"id": n["id"], "parent": n["parent"], "c_change": {k: v for k, v in n["c_change"]},
                  "c": {k: v for k, v in n["c"]} if "c" in n else None} for n in tree_json["nodes"]]
        return cls(nodes, tree_json["nu"])
