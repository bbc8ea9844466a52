"""String world -> interned columnar tables (``cae_objects``).

This is the Python twin of what the Go shim does once per autoscaler tick: walk
the snapshot's NodeInfos, the template NodeInfos and the pending pod groups, intern every string,
and lay the result out as the CSR tables ``include/caengine.h`` declares.  It evaluates no
predicate.  Two layers:

* :class:`TableBuilder` — integer-level, de-duplicating table construction (used directly by the
  synthetic generator for 10^5..10^6 pods);
* :class:`Encoder`      — interns ``objects.py`` dataclasses on top of it.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .objects import (LABEL_HOSTNAME, TAINT_NODE_UNSCHEDULABLE, LabelSelector, Namespace, Node, NodeInfo, Pod,
                      PodEquivalenceGroup, Requirement)

MAX_RES = capi.CONST["CAE_MAX_RES"]

_OPS = {"In": 0, "Equals": 0, "=": 0, "==": 0, "NotIn": 1, "!=": 1, "Exists": 2,
        "DoesNotExist": 3, "Gt": 4, "Lt": 5}
_TOL_OPS = {"": 0, "Equal": 0, "Exists": 1, "Lt": 2, "Gt": 3}
_EFFECTS = {"": 0, "NoSchedule": 1, "PreferNoSchedule": 2, "NoExecute": 3}
_PROTOS = {"": 0, "TCP": 0, "UDP": 1, "SCTP": 2}
_POLICY = {"Ignore": 0, "Honor": 1}


class Unsupported(Exception):
    """Input the engine refuses (SURVEY §7 hard part 7): the caller must use the stock Go path."""


class _Interner:
    def __init__(self) -> None:
        self.ids: Dict[object, int] = {}
        self.items: List[object] = []

    def __call__(self, x) -> int:
        i = self.ids.get(x)
        if i is None:
            i = len(self.items)
            self.ids[x] = i
            self.items.append(x)
        return i

    def __len__(self) -> int:
        return len(self.items)


class _Csr:
    """A de-duplicating list-of-lists table; list 0 is the empty list."""

    def __init__(self, ncols: int) -> None:
        self.ncols = ncols
        self.ids: Dict[tuple, int] = {(): 0}
        self.off: List[int] = [0, 0]
        self.cols: List[List[int]] = [[] for _ in range(ncols)]

    def add(self, rows: Sequence[tuple], dedupe: bool = True) -> int:
        key = tuple(rows)
        if dedupe:
            i = self.ids.get(key)
            if i is not None:
                return i
        for r in rows:
            for c in range(self.ncols):
                self.cols[c].append(int(r[c]))
        self.off.append(self.off[-1] + len(rows))
        i = len(self.off) - 2
        if dedupe:
            self.ids[key] = i
        return i

    @property
    def n(self) -> int:
        return len(self.off) - 1


def _i32(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.int32)
    return a if a.size else np.zeros(1, np.int32)[:0].copy()


class TableBuilder:
    """Integer-level construction of every table of ``cae_objects``."""

    def __init__(self, num_res: int = 3) -> None:
        self.num_res = num_res
        self.hostname_key = -1
        self.unschedulable_taint_key = -1
        self.value_is_int: List[int] = []
        self.value_int: List[int] = []
        self.ns_labelset: List[int] = []
        self.ns_exists: List[int] = []
        self.labelsets = _Csr(2)
        self.reqs_key: List[int] = []
        self.reqs_op: List[int] = []
        self.reqs_val_off: List[int] = [0]
        self.reqs_vals: List[int] = []
        self.sel_ids: Dict[tuple, int] = {}
        self.sel_kind: List[int] = []
        self.sel_req_off: List[int] = [0]
        self.naff_ids: Dict[tuple, int] = {}
        self.naff_nodesel: List[int] = []
        self.naff_has_required: List[int] = []
        self.naff_term_off: List[int] = [0]
        self.term_expr_sel: List[int] = []
        self.term_field_off: List[int] = [0]
        self.field_op: List[int] = []
        self.field_node_name: List[int] = []
        self.tols = _Csr(4)
        self.taints = _Csr(3)
        self.ports = _Csr(3)
        self.pts = _Csr(6)
        # affinity lists: list of term ids; terms are rows with their own namespace CSR
        self.aff_ids: Dict[tuple, int] = {(): 0}
        self.aff_off: List[int] = [0, 0]
        self.aterm_selector: List[int] = []
        self.aterm_key: List[int] = []
        self.aterm_ns_off: List[int] = [0]
        self.aterm_ns: List[int] = []
        self.aterm_ns_selector: List[int] = []
        self.ps_ids: Dict[tuple, int] = {}
        self.ps_rows: List[tuple] = []
        # nodes
        self.node_rows: List[tuple] = []
        self.node_alloc: List[List[int]] = []
        self.node_pod_off: List[int] = [0]
        self.node_pod_spec: List[int] = []
        self.num_cluster_nodes = 0
        self.num_templates = 0
        # pending
        self.group_off: List[int] = [0]
        self.pend_spec_chunks: List[np.ndarray] = []
        self._nothing_sel = None

    # ---- dictionaries -------------------------------------------------------------------
    def declare_value(self, vid: int, text: Optional[str]) -> None:
        while len(self.value_is_int) <= vid:
            self.value_is_int.append(0)
            self.value_int.append(0)
        if text is not None:
            ok, v = _parse_int64(text)
            self.value_is_int[vid] = int(ok)
            self.value_int[vid] = v

    def declare_namespace(self, nsid: int, labelset: int = 0, exists: bool = False) -> None:
        while len(self.ns_labelset) <= nsid:
            self.ns_labelset.append(0)
            self.ns_exists.append(0)
        self.ns_labelset[nsid] = labelset
        self.ns_exists[nsid] = int(exists)

    # ---- tables ---------------------------------------------------------------------------
    def labelset(self, pairs: Iterable[Tuple[int, int]]) -> int:
        return self.labelsets.add(sorted(pairs))

    def selector(self, reqs: Optional[Sequence[Tuple[int, int, Tuple[int, ...]]]]) -> int:
        """reqs = None -> Nothing; [] -> Everything; else AND of (key, op, values)."""
        key = ("nothing",) if reqs is None else tuple((k, o, tuple(v)) for k, o, v in reqs)
        i = self.sel_ids.get(key)
        if i is not None:
            return i
        i = len(self.sel_kind)
        self.sel_ids[key] = i
        self.sel_kind.append(0 if reqs is None else 1)
        for k, o, vals in (reqs or []):
            self.reqs_key.append(k)
            self.reqs_op.append(o)
            self.reqs_vals.extend(vals)
            self.reqs_val_off.append(len(self.reqs_vals))
        self.sel_req_off.append(len(self.reqs_key))
        return i

    def nothing_selector(self) -> int:
        return self.selector(None)

    def node_affinity(self, nodesel: int, has_required: bool,
                      terms: Sequence[Tuple[int, Sequence[Tuple[int, int]]]]) -> int:
        """terms = [(expr_selector_or_-1, [(field_op, node_name_id), ...]), ...] (empty terms dropped)."""
        key = (nodesel, bool(has_required), tuple((e, tuple(f)) for e, f in terms))
        i = self.naff_ids.get(key)
        if i is not None:
            return i
        i = len(self.naff_nodesel)
        self.naff_ids[key] = i
        self.naff_nodesel.append(nodesel)
        self.naff_has_required.append(int(has_required))
        for e, fields in terms:
            self.term_expr_sel.append(e)
            for op, nm in fields:
                self.field_op.append(op)
                self.field_node_name.append(nm)
            self.term_field_off.append(len(self.field_op))
        self.naff_term_off.append(len(self.term_expr_sel))
        return i

    def toleration_list(self, tols: Sequence[Tuple[int, int, int, int]]) -> int:
        return self.tols.add(list(tols))

    def taint_list(self, taints: Sequence[Tuple[int, int, int]]) -> int:
        return self.taints.add(list(taints))

    def port_list(self, ports: Sequence[Tuple[int, int, int]]) -> int:
        return self.ports.add(list(ports))

    def pts_list(self, cons: Sequence[Tuple[int, int, int, int, int, int]]) -> int:
        """(max_skew, key, selector, min_domains, node_affinity_policy, node_taints_policy)"""
        return self.pts.add(list(cons))

    def affinity_list(self, terms: Sequence[Tuple[int, int, Tuple[int, ...], int]]) -> int:
        """terms = [(selector, topology_key, namespaces, ns_selector)]"""
        key = tuple((s, k, tuple(ns), nss) for s, k, ns, nss in terms)
        i = self.aff_ids.get(key)
        if i is not None:
            return i
        for s, k, ns, nss in terms:
            self.aterm_selector.append(s)
            self.aterm_key.append(k)
            self.aterm_ns.extend(ns)
            self.aterm_ns_off.append(len(self.aterm_ns))
            self.aterm_ns_selector.append(nss)
        self.aff_off.append(len(self.aterm_selector))
        i = len(self.aff_off) - 2
        self.aff_ids[key] = i
        return i

    def podspec(self, namespace: int, labelset: int, req: Sequence[int], tol_list: int = 0,
                naff: int = -1, node_name: int = -1, port_list: int = 0, pts_list: int = 0,
                aff_list: int = 0, anti_list: int = 0, terminating: bool = False,
                hostname_spread: Optional[bool] = None) -> int:
        req = tuple(int(x) for x in req) + (0,) * (MAX_RES - len(req))
        if hostname_spread is None:  # default: derived from the hard constraints
            hostname_spread = any(self.pts.cols[1][i] == self.hostname_key
                                  for i in range(self.pts.off[pts_list], self.pts.off[pts_list + 1]))
        key = (namespace, labelset, req, tol_list, naff, node_name, port_list, pts_list, aff_list,
               anti_list, bool(terminating), bool(hostname_spread))
        i = self.ps_ids.get(key)
        if i is None:
            i = len(self.ps_rows)
            self.ps_ids[key] = i
            self.ps_rows.append(key)
        return i

    def _node(self, name: int, labelset: int, taint_list: int, unschedulable: bool,
              alloc: Sequence[int], allowed_pods: int, cap_cpu: int, cap_mem: int,
              has_alloc_cpu: bool, has_alloc_mem: bool, pod_specs: Sequence[int]) -> int:
        self.node_rows.append((name, labelset, taint_list, int(unschedulable), allowed_pods, cap_cpu,
                               cap_mem, int(has_alloc_cpu), int(has_alloc_mem)))
        self.node_alloc.append(list(alloc) + [0] * (MAX_RES - len(alloc)))
        self.node_pod_spec.extend(pod_specs)
        self.node_pod_off.append(len(self.node_pod_spec))
        return len(self.node_rows) - 1

    def cluster_node(self, *a, **kw) -> int:
        if self.num_templates:
            raise ValueError("cluster nodes must be added before templates")
        self.num_cluster_nodes += 1
        return self._node(*a, **kw)

    def template(self, *a, **kw) -> int:
        self.num_templates += 1
        return self._node(*a, **kw) - self.num_cluster_nodes

    def group(self, spec_ids: Sequence[int]) -> int:
        arr = np.asarray(spec_ids, dtype=np.int32)
        self.pend_spec_chunks.append(arr)
        self.group_off.append(self.group_off[-1] + len(arr))
        return len(self.group_off) - 2

    # ---- finish ---------------------------------------------------------------------------
    def finish(self) -> "EncodedObjects":
        return EncodedObjects(self)


def _parse_int64(s: str) -> Tuple[bool, int]:
    """strconv.ParseInt(s, 10, 64): optional sign, decimal digits only (underscores not allowed with base 10)."""
    t = s
    if t[:1] in "+-":
        t = t[1:]
    if not t or not t.isascii() or not t.isdigit():
        return False, 0
    v = int(s)
    if v < -(1 << 63) or v > (1 << 63) - 1:
        return False, 0
    return True, v


class EncodedObjects:
    """Owns the numpy arrays behind one ``cae_objects`` struct."""

    def __init__(self, b: TableBuilder) -> None:
        a: Dict[str, np.ndarray] = {}
        nvals = max(len(b.value_is_int), 1)
        a["value_is_int"] = np.zeros(nvals, np.uint8)
        a["value_int"] = np.zeros(nvals, np.int64)
        a["value_is_int"][:len(b.value_is_int)] = b.value_is_int
        a["value_int"][:len(b.value_int)] = b.value_int
        nns = max(len(b.ns_labelset), 1)
        a["ns_labelset"] = np.zeros(nns, np.int32)
        a["ns_exists"] = np.zeros(nns, np.uint8)
        a["ns_labelset"][:len(b.ns_labelset)] = b.ns_labelset
        a["ns_exists"][:len(b.ns_exists)] = b.ns_exists
        a["ls_off"] = _i32(b.labelsets.off)
        a["ls_key"] = _i32(b.labelsets.cols[0])
        a["ls_val"] = _i32(b.labelsets.cols[1])
        a["req_key"] = _i32(b.reqs_key)
        a["req_op"] = _i32(b.reqs_op)
        a["req_val_off"] = _i32(b.reqs_val_off)
        a["req_vals"] = _i32(b.reqs_vals)
        a["sel_kind"] = _i32(b.sel_kind)
        a["sel_req_off"] = _i32(b.sel_req_off)
        a["naff_nodesel"] = _i32(b.naff_nodesel)
        a["naff_has_required"] = np.asarray(b.naff_has_required, np.uint8)
        a["naff_term_off"] = _i32(b.naff_term_off)
        a["term_expr_sel"] = _i32(b.term_expr_sel)
        a["term_field_off"] = _i32(b.term_field_off)
        a["field_op"] = _i32(b.field_op)
        a["field_node_name"] = _i32(b.field_node_name)
        a["tol_off"] = _i32(b.tols.off)
        for i, nm in enumerate(("tol_key", "tol_op", "tol_val", "tol_effect")):
            a[nm] = _i32(b.tols.cols[i])
        a["taint_off"] = _i32(b.taints.off)
        for i, nm in enumerate(("taint_key", "taint_val", "taint_effect")):
            a[nm] = _i32(b.taints.cols[i])
        a["port_off"] = _i32(b.ports.off)
        for i, nm in enumerate(("port_ip", "port_proto", "port_num")):
            a[nm] = _i32(b.ports.cols[i])
        a["pts_off"] = _i32(b.pts.off)
        for i, nm in enumerate(("pts_max_skew", "pts_key", "pts_selector", "pts_min_domains",
                                "pts_node_affinity_policy", "pts_node_taints_policy")):
            a[nm] = _i32(b.pts.cols[i])
        a["aff_off"] = _i32(b.aff_off)
        a["aterm_selector"] = _i32(b.aterm_selector)
        a["aterm_key"] = _i32(b.aterm_key)
        a["aterm_ns_off"] = _i32(b.aterm_ns_off)
        a["aterm_ns"] = _i32(b.aterm_ns)
        a["aterm_ns_selector"] = _i32(b.aterm_ns_selector)
        nps = len(b.ps_rows)
        cols = list(zip(*b.ps_rows)) if nps else [[] for _ in range(12)]
        a["ps_namespace"] = _i32(cols[0])
        a["ps_labelset"] = _i32(cols[1])
        a["ps_req"] = np.ascontiguousarray(np.asarray(cols[2], dtype=np.int64).reshape(nps, MAX_RES))
        a["ps_tol_list"] = _i32(cols[3])
        a["ps_naff"] = _i32(cols[4])
        a["ps_node_name"] = _i32(cols[5])
        a["ps_port_list"] = _i32(cols[6])
        a["ps_pts_list"] = _i32(cols[7])
        a["ps_aff_list"] = _i32(cols[8])
        a["ps_anti_list"] = _i32(cols[9])
        a["ps_terminating"] = np.asarray(cols[10], dtype=np.uint8)
        a["ps_hostname_spread"] = np.asarray(cols[11], dtype=np.uint8)
        nn = len(b.node_rows)
        ncols = list(zip(*b.node_rows)) if nn else [[] for _ in range(9)]
        a["node_name"] = _i32(ncols[0])
        a["node_labelset"] = _i32(ncols[1])
        a["node_taint_list"] = _i32(ncols[2])
        a["node_unschedulable"] = np.asarray(ncols[3], dtype=np.uint8)
        a["node_allowed_pods"] = _i32(ncols[4])
        a["node_cap_cpu"] = np.asarray(ncols[5], dtype=np.int64)
        a["node_cap_mem"] = np.asarray(ncols[6], dtype=np.int64)
        a["node_has_alloc_cpu"] = np.asarray(ncols[7], dtype=np.uint8)
        a["node_has_alloc_mem"] = np.asarray(ncols[8], dtype=np.uint8)
        a["node_alloc"] = np.ascontiguousarray(np.asarray(b.node_alloc, dtype=np.int64).reshape(nn, MAX_RES))
        a["node_pod_off"] = _i32(b.node_pod_off)
        a["node_pod_spec"] = _i32(b.node_pod_spec)
        a["group_off"] = _i32(b.group_off)
        a["pend_spec"] = (np.ascontiguousarray(np.concatenate(b.pend_spec_chunks).astype(np.int32))
                          if b.pend_spec_chunks else np.zeros(0, np.int32))
        self.arrays = a
        s = capi.cae_objects()
        s.abi_version = capi.CONST["CAE_ABI_VERSION"]
        s.num_res = b.num_res
        s.num_values = nvals
        s.hostname_key = b.hostname_key
        s.unschedulable_taint_key = b.unschedulable_taint_key
        s.num_namespaces = nns
        s.num_labelsets = b.labelsets.n
        s.num_reqs = len(b.reqs_key)
        s.num_selectors = len(b.sel_kind)
        s.num_naff = len(b.naff_nodesel)
        s.num_naff_terms = len(b.term_expr_sel)
        s.num_tol_lists = b.tols.n
        s.num_taint_lists = b.taints.n
        s.num_port_lists = b.ports.n
        s.num_pts_lists = b.pts.n
        s.num_aff_lists = len(b.aff_off) - 1
        s.num_aterms = len(b.aterm_selector)
        s.num_podspecs = nps
        s.num_cluster_nodes = b.num_cluster_nodes
        s.num_templates = b.num_templates
        s.num_groups = len(b.group_off) - 1
        s.num_pending = int(b.group_off[-1])
        for name, ctype in capi.cae_objects._fields_:
            if name in a:
                arr = a[name]
                setattr(s, name, arr.ctypes.data_as(ctype))
        missing = [n for n, t in capi.cae_objects._fields_
                   if n not in a and hasattr(t, "contents")]
        if missing:
            raise RuntimeError("encoder does not fill ABI fields: %s" % missing)
        self.struct = s

    def slice_pods(self, p_begin: int, p_end: int) -> "EncodedObjects":
        """The same snapshot with only the pending pods [p_begin, p_end) (groups clipped to the range; every other table is
        shared): what rank r of a pod-sharded dense pass uploads (CAE_CFG_PODS_PRESHARDED)."""
        import copy
        out = copy.copy(self)
        out.arrays = dict(self.arrays)
        go = np.clip(self.arrays["group_off"], p_begin, p_end) - p_begin
        out.arrays["group_off"] = np.ascontiguousarray(go.astype(np.int32))
        out.arrays["pend_spec"] = np.ascontiguousarray(self.arrays["pend_spec"][p_begin:p_end])
        s = capi.cae_objects()
        C.memmove(C.byref(s), C.byref(self.struct), C.sizeof(s))
        s.num_pending = p_end - p_begin
        for name, ctype in capi.cae_objects._fields_:
            if name in ("group_off", "pend_spec"):
                setattr(s, name, out.arrays[name].ctypes.data_as(ctype))
        out.struct = s
        return out

    # convenience
    @property
    def P(self) -> int:
        return self.struct.num_pending

    @property
    def T(self) -> int:
        return self.struct.num_templates

    @property
    def E(self) -> int:
        return self.struct.num_groups

    def ptr(self):
        return C.byref(self.struct)


# ----------------------------------------------------------------------------------------------
class Encoder:
    """Interns objects.py dataclasses into a TableBuilder (the Go shim's job in production)."""

    FIXED_RES = {"cpu": 0, "memory": 1, "ephemeral-storage": 2}

    def __init__(self) -> None:
        self.keys = _Interner()
        self.values = _Interner()
        self.namespaces = _Interner()
        self.node_names = _Interner()
        self.ips = _Interner()
        self.ips("0.0.0.0")
        self.resources = _Interner()
        for r in ("cpu", "memory", "ephemeral-storage"):
            self.resources(r)
        self.b = TableBuilder()
        self._podspec_cache: Dict[int, int] = {}
        self._ns_objects: Dict[str, Namespace] = {}

    # ---- interning helpers --------------------------------------------------------------
    def _val(self, v: str) -> int:
        n = len(self.values)
        i = self.values(v)
        if i == n:
            self.b.declare_value(i, v)
        return i

    def _key(self, k: str) -> int:
        i = self.keys(k)
        if k == LABEL_HOSTNAME:
            self.b.hostname_key = i
        elif k == TAINT_NODE_UNSCHEDULABLE:
            self.b.unschedulable_taint_key = i
        return i

    def _ns(self, ns: str) -> int:
        n = len(self.namespaces)
        i = self.namespaces(ns)
        if i == n:
            obj = self._ns_objects.get(ns)
            if obj is not None:
                self.b.declare_namespace(i, self._labelset(obj.labels), True)
            else:
                self.b.declare_namespace(i, 0, False)
        return i

    def _labelset(self, labels: Dict[str, str]) -> int:
        return self.b.labelset((self._key(k), self._val(v)) for k, v in labels.items())

    def _selector(self, sel: Optional[LabelSelector], extra: Optional[Dict[str, str]] = None) -> int:
        """metav1.LabelSelectorAsSelector; `extra` = matchLabelKeys merge (common.go:96-106,131-143)."""
        if sel is None:
            # mergeLabelSetWithSelector on Nothing: Requirements() of Nothing is not ok -> returns s
            return self.b.nothing_selector()
        reqs: List[Tuple[int, int, Tuple[int, ...]]] = []
        for k, v in sorted((extra or {}).items()):
            reqs.append((self._key(k), 0, (self._val(v),)))
        for k, v in sorted(sel.match_labels.items()):
            reqs.append((self._key(k), 0, (self._val(v),)))
        for r in sel.match_expressions:
            reqs.append(self._req(r))
        return self.b.selector(reqs)

    def _req(self, r: Requirement) -> Tuple[int, int, Tuple[int, ...]]:
        if r.operator not in _OPS:
            raise Unsupported("selector operator %r" % r.operator)
        return (self._key(r.key), _OPS[r.operator], tuple(self._val(v) for v in r.values))

    def _resource_vec(self, rl: Dict[str, int]) -> List[int]:
        vec = [0] * MAX_RES
        for name, amt in rl.items():
            if name == "pods":
                continue
            i = self.resources(name)
            if i >= MAX_RES:
                raise Unsupported("more than %d resource dimensions" % MAX_RES)
            vec[i] = int(amt)
        self.b.num_res = max(self.b.num_res, len(self.resources))
        return vec

    # ---- objects --------------------------------------------------------------------------
    def add_namespace(self, ns: Namespace) -> None:
        self._ns_objects[ns.name] = ns
        if ns.name in self.namespaces.ids:
            self.b.declare_namespace(self.namespaces.ids[ns.name], self._labelset(ns.labels), True)

    def podspec(self, pod: Pod, resident: bool = False) -> int:
        """resident: a pod already running on a cluster node.  The volume / DRA filters (VolumeRestrictions, VolumeBinding,
        VolumeZone, NodeVolumeLimits, DynamicResources) only ever reject an INCOMING pod that carries volumes or claims, so
        residents with volumes are harmless as long as no pending (or template DaemonSet) pod has any — those are refused."""
        if pod.has_volumes_or_claims and not resident:
            raise Unsupported("pod %s/%s uses volumes or resource claims" % (pod.namespace, pod.name))
        cached = self._podspec_cache.get(id(pod))
        if cached is not None:
            return cached
        b = self.b
        ns = self._ns(pod.namespace)
        tols = b.toleration_list([
            (self._key(t.key) if t.key else -1, _TOL_OPS.get(t.operator, 4),
             self._val(t.value) if t.value else -1, _EFFECTS[t.effect]) for t in pod.tolerations])
        naff = -1
        if pod.node_selector or pod.node_affinity_terms is not None:
            nodesel = -1
            if pod.node_selector:
                nodesel = b.selector([(self._key(k), 0, (self._val(v),))
                                      for k, v in sorted(pod.node_selector.items())])
            terms = []
            for t in (pod.node_affinity_terms or []):
                # empty terms are kept: they select nothing in Filter (nodeaffinity.go:60-66) but make
                # NodeAffinity.PreFilter return "all nodes" (node_affinity.go:176-196)
                expr = -1
                if t.match_expressions:
                    expr = b.selector([self._req(r) for r in t.match_expressions])
                flds = []
                for f in t.match_fields:
                    if f.key != "metadata.name" or f.operator not in ("In", "NotIn") or len(f.values) != 1:
                        raise Unsupported("matchFields other than metadata.name In/NotIn [one value]")
                    flds.append((_OPS[f.operator], self.node_names(f.values[0])))
                terms.append((expr, flds))
            naff = b.node_affinity(nodesel, pod.node_affinity_terms is not None, terms)
        ports = b.port_list([(self.ips(p.host_ip or "0.0.0.0"), _PROTOS[p.protocol], p.host_port)
                             for p in pod.host_ports if p.host_port > 0])
        cons = []
        for c in pod.topology_spread:
            if c.when_unsatisfiable != "DoNotSchedule":
                continue  # plugin.go:273-296 keeps only the hard constraints
            extra = {k: pod.labels[k] for k in c.match_label_keys if k in pod.labels}
            cons.append((c.max_skew, self._key(c.topology_key), self._selector(c.label_selector, extra),
                         1 if c.min_domains is None else c.min_domains,
                         _POLICY[c.node_affinity_policy or "Honor"],
                         _POLICY[c.node_taints_policy or "Ignore"]))
        pts = b.pts_list(cons)

        def aff(terms) -> int:
            rows = []
            for t in terms:
                nss = list(t.namespaces)
                if not nss and t.namespace_selector is None:
                    nss = [pod.namespace]  # types.go:436-444
                rows.append((self._selector(t.label_selector), self._key(t.topology_key),
                             tuple(sorted(self._ns(n) for n in nss)),
                             self._selector(t.namespace_selector)))
            return b.affinity_list(rows)

        sid = b.podspec(ns, self._labelset(pod.labels), self._resource_vec(pod.requests), tols, naff,
                        self.node_names(pod.node_name) if pod.node_name else -1, ports, pts,
                        aff(pod.pod_affinity), aff(pod.pod_anti_affinity), pod.terminating,
                        any(c.topology_key == LABEL_HOSTNAME for c in pod.topology_spread))
        self._podspec_cache[id(pod)] = sid
        return sid

    def _node_args(self, ni: NodeInfo, resident: bool = False):
        n = ni.node
        taints = self.b.taint_list([(self._key(t.key), self._val(t.value) if t.value else -1,
                                     _EFFECTS[t.effect]) for t in n.taints])
        return dict(name=self.node_names(n.name), labelset=self._labelset(n.labels), taint_list=taints,
                    unschedulable=n.unschedulable, alloc=self._resource_vec(n.allocatable),
                    allowed_pods=int(n.allocatable.get("pods", 0)),
                    cap_cpu=int(n.capacity.get("cpu", 0)), cap_mem=int(n.capacity.get("memory", 0)),
                    has_alloc_cpu="cpu" in n.allocatable, has_alloc_mem="memory" in n.allocatable,
                    pod_specs=[self.podspec(p, resident) for p in ni.pods])

    def add_cluster_node(self, ni: NodeInfo) -> int:
        return self.b.cluster_node(**self._node_args(ni, resident=True))

    def add_template(self, ni: NodeInfo) -> int:
        return self.b.template(**self._node_args(ni))

    def add_group(self, g: PodEquivalenceGroup) -> int:
        return self.b.group([self.podspec(p) for p in g.pods])

    def finish(self) -> EncodedObjects:
        self._key(LABEL_HOSTNAME)
        self._key(TAINT_NODE_UNSCHEDULABLE)
        self.b.num_res = max(3, len(self.resources))
        return self.b.finish()


def encode(cluster: Sequence[NodeInfo], templates: Sequence[NodeInfo],
           groups: Sequence[PodEquivalenceGroup],
           namespaces: Sequence[Namespace] = ()) -> EncodedObjects:
    enc = Encoder()
    for ns in namespaces:
        enc.add_namespace(ns)
    for ni in cluster:
        enc.add_cluster_node(ni)
    for ni in templates:
        enc.add_template(ni)
    for g in groups:
        enc.add_group(g)
    return enc.finish()
