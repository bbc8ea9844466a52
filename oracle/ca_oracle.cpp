/*
 * ca_oracle.cpp — CPU restatement of the Cluster Autoscaler scale-up simulation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (kubernetes-autoscaler_b200/, libcaengine.so)
 * links, imports or executes this file; it is the checker the parity tests and bench.py's
 * cpu_baseline leg compare the CUDA engine against.  The reference (100 % Go, needs Go 1.25) cannot
 * be built in this image, so this is a "port": every function cites the reference file:line it
 * restates (paths relative to /root/reference/cluster-autoscaler; K8S = vendor/k8s.io/kubernetes/
 * pkg/scheduler).  Parity is PINNED against the reference's own known-answer tests
 * (estimator/binpacking_estimator_test.go:90-296 etc.) in tests/test_oracle_kat.py.
 * Parity UNPINNED (no reference test at this boundary): tie order of equal DecreasingPodOrderer
 * scores (Go's sort.Slice is unstable; we use a stable sort), node list order (Go map order in the
 * reference; insertion order here), lastIndex (reset at the start of every Estimate, i.e. every
 * Estimate behaves as on a fresh PredicateSnapshot, as in every reference test).
 *
 * Faithful to the reference's complexity where it matters for a CPU baseline: PodTopologySpread and
 * InterPodAffinity PreFilter rescan all nodes/pods for every SchedulePod, exactly like
 * K8S/framework/plugins/podtopologyspread/filtering.go:237-311 and interpodaffinity/filtering.go:204-271.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../include/caengine.h"

namespace {

constexpr int R = CAE_MAX_RES;
using i64 = int64_t;

struct Port { int ip, proto, port; };

/* framework.NodeInfo (K8S/framework/types.go:165-208) reduced to what the Filter plugins read. */
struct NodeState {
  int name = -1;         /* node-name id; fresh nodes get unique ids < -1 */
  int labelset = 0;
  int hostname_val = -1; /* >= -1: unused; < -1: sanitized node, hostname label := fresh value (node_info_utils.go:127) */
  bool sanitized = false;
  bool unschedulable = false;
  int taint_list = 0;
  i64 alloc[R] = {0};
  int allowed_pods = 0;
  i64 requested[R] = {0}; /* NodeInfo.Requested, update() types.go:423-442 */
  std::vector<int> pods;  /* podspec ids, NodeInfo.Pods */
  std::vector<Port> used_ports;
  bool is_new = false;    /* estimationState.newNodeNames */
  bool has_sched = false; /* estimationState.newNodesWithPods */
};

struct Ctx {
  const cae_objects* o;
  explicit Ctx(const cae_objects* oo) : o(oo) {}

  /* ---- labels ---- */
  /* node label lookup honouring the sanitized hostname label */
  bool nodeLabel(const NodeState& n, int key, int* val) const {
    if (n.sanitized && key == o->hostname_key && key >= 0) { *val = n.hostname_val; return true; }
    for (int i = o->ls_off[n.labelset]; i < o->ls_off[n.labelset + 1]; ++i)
      if (o->ls_key[i] == key) { *val = o->ls_val[i]; return true; }
    return false;
  }
  bool valueInt(int v, i64* out) const {
    if (v < 0 || v >= o->num_values || !o->value_is_int[v]) return false;
    *out = o->value_int[v];
    return true;
  }
  /* Requirement.Matches (apimachinery/pkg/labels/selector.go:247-294).
   * lookup(key,&val) abstracts labels.Set vs sanitized node labels. */
  template <class Lookup>
  bool reqMatches(int r, Lookup&& lookup) const {
    int key = o->req_key[r], op = o->req_op[r];
    int vb = o->req_val_off[r], ve = o->req_val_off[r + 1];
    int val;
    bool has = lookup(key, &val);
    auto hasValue = [&](int v) { for (int i = vb; i < ve; ++i) if (o->req_vals[i] == v) return true; return false; };
    switch (op) {
      case CAE_OP_IN: return has && hasValue(val);
      case CAE_OP_NOT_IN: return !has || !hasValue(val);
      case CAE_OP_EXISTS: return has;
      case CAE_OP_DOES_NOT_EXIST: return !has;
      case CAE_OP_GT: case CAE_OP_LT: {
        if (!has) return false;
        i64 lv, rv;
        if (!valueInt(val, &lv)) return false;
        if (ve - vb != 1) return false;
        if (!valueInt(o->req_vals[vb], &rv)) return false;
        return (op == CAE_OP_GT && lv > rv) || (op == CAE_OP_LT && lv < rv);
      }
      default: return false;
    }
  }
  template <class Lookup>
  bool selMatches(int s, Lookup&& lookup) const { /* internalSelector.Matches / nothingSelector */
    if (o->sel_kind[s] == CAE_SEL_NOTHING) return false;
    for (int r = o->sel_req_off[s]; r < o->sel_req_off[s + 1]; ++r)
      if (!reqMatches(r, lookup)) return false;
    return true;
  }
  bool selEmpty(int s) const { /* Selector.Empty(): Everything only (selector.go:96-118) */
    return o->sel_kind[s] == CAE_SEL_REQS && o->sel_req_off[s] == o->sel_req_off[s + 1];
  }
  bool selMatchesLabelset(int s, int ls) const {
    return selMatches(s, [&](int key, int* val) {
      for (int i = o->ls_off[ls]; i < o->ls_off[ls + 1]; ++i)
        if (o->ls_key[i] == key) { *val = o->ls_val[i]; return true; }
      return false;
    });
  }
  bool selMatchesNode(int s, const NodeState& n) const {
    return selMatches(s, [&](int key, int* val) { return nodeLabel(n, key, val); });
  }

  /* RequiredNodeAffinity.Match (component-helpers/scheduling/corev1/nodeaffinity/nodeaffinity.go:323-334)
   * with LazyErrorNodeSelector.Match (:85-106) and nodeSelectorTerm.match (:203-214). */
  bool naffMatch(int a, const NodeState& n) const {
    if (a < 0) return true;
    if (o->naff_nodesel[a] >= 0 && !selMatchesNode(o->naff_nodesel[a], n)) return false;
    if (!o->naff_has_required[a]) return true;
    for (int t = o->naff_term_off[a]; t < o->naff_term_off[a + 1]; ++t) {
      int fb = o->term_field_off[t], fe = o->term_field_off[t + 1];
      if (o->term_expr_sel[t] < 0 && fb == fe) continue; /* empty term selects no objects (:60-66) */
      if (o->term_expr_sel[t] >= 0 && !selMatchesNode(o->term_expr_sel[t], n)) continue;
      bool ok = true;
      for (int f = fb; f < fe && ok; ++f) { /* fields.OneTermEqual/NotEqualSelector on metadata.name */
        bool eq = (o->field_node_name[f] == n.name);
        ok = (o->field_op[f] == CAE_OP_IN) ? eq : !eq;
      }
      if (ok) return true;
    }
    return false;
  }

  /* Toleration.ToleratesTaint (vendor/k8s.io/api/core/v1/toleration.go:52-77), comparison ops gate off */
  bool tolerates(int ti, int tkey, int tval, int teffect) const {
    if (o->tol_effect[ti] != CAE_EFFECT_NONE && o->tol_effect[ti] != teffect) return false;
    if (o->tol_key[ti] >= 0 && o->tol_key[ti] != tkey) return false;
    switch (o->tol_op[ti]) {
      case CAE_TOL_EQUAL: return o->tol_val[ti] == tval;
      case CAE_TOL_EXISTS: return true;
      default: return false; /* Lt/Gt: TaintTolerationComparisonOperators off; unknown op */
    }
  }
  bool tolerationsTolerate(int tol_list, int tkey, int tval, int teffect) const {
    for (int i = o->tol_off[tol_list]; i < o->tol_off[tol_list + 1]; ++i)
      if (tolerates(i, tkey, tval, teffect)) return true;
    return false;
  }
  /* FindMatchingUntoleratedTaint with DoNotScheduleTaintsFilterFunc
   * (component-helpers/scheduling/corev1/helpers.go:79-87, K8S/framework/plugins/helper/taint.go:23-28) */
  bool hasUntoleratedTaint(int taint_list, int tol_list) const {
    for (int i = o->taint_off[taint_list]; i < o->taint_off[taint_list + 1]; ++i) {
      int eff = o->taint_effect[i];
      if (eff != CAE_EFFECT_NO_SCHEDULE && eff != CAE_EFFECT_NO_EXECUTE) continue;
      if (!tolerationsTolerate(tol_list, o->taint_key[i], o->taint_val[i], eff)) return true;
    }
    return false;
  }

  /* AffinityTerm.Matches (kube-scheduler/framework/types.go:387-392) */
  bool atermNsHas(int t, int ns) const {
    for (int i = o->aterm_ns_off[t]; i < o->aterm_ns_off[t + 1]; ++i) if (o->aterm_ns[i] == ns) return true;
    return false;
  }
  /* existing pod's term vs the incoming pod, with the incoming pod's namespace labels
   * (interpodaffinity/filtering.go:213, plugin.go:161-170: missing Namespace object -> empty labels) */
  bool atermMatchesWithNsLabels(int t, int pod_spec) const {
    int ns = o->ps_namespace[pod_spec];
    int nsls = o->ns_exists[ns] ? o->ns_labelset[ns] : 0;
    if (atermNsHas(t, ns) || selMatchesLabelset(o->aterm_ns_selector[t], nsls))
      return selMatchesLabelset(o->aterm_selector[t], o->ps_labelset[pod_spec]);
    return false;
  }
  /* incoming pod's term vs an existing pod: namespaces merged from the lister, nsLabels = nil
   * (plugin.go:144-157; the by-value `at` means NamespaceSelector itself is NOT replaced, so a
   * non-empty selector is still evaluated against the empty label set — kept as in the reference). */
  bool incomingTermMatches(int t, int other_spec) const {
    int ns = o->ps_namespace[other_spec];
    bool nsok = atermNsHas(t, ns);
    int nss = o->aterm_ns_selector[t];
    if (!nsok && !selEmpty(nss) && o->ns_exists[ns] && selMatchesLabelset(nss, o->ns_labelset[ns])) nsok = true; /* merged */
    if (!nsok && selMatchesLabelset(nss, 0)) nsok = true; /* NamespaceSelector.Matches(nil) */
    if (!nsok) return false;
    return selMatchesLabelset(o->aterm_selector[t], o->ps_labelset[other_spec]);
  }
};

/* HostPortInfo.CheckConflict (kube-scheduler/framework/types.go:599-628); ip id 0 == 0.0.0.0 */
bool portConflict(const std::vector<Port>& used, const Port& w) {
  if (w.port <= 0) return false;
  for (const Port& u : used) {
    if (u.proto != w.proto || u.port != w.port) continue;
    if (w.ip == 0 || u.ip == 0 || u.ip == w.ip) return true;
  }
  return false;
}

struct PtsCon { int max_skew, key, sel, min_domains, aff_policy, taint_policy; };

/* CycleState after RunPreFilterPlugins (K8S/framework/runtime/framework.go:862-923) */
struct PreFilter {
  bool failed = false;          /* NodeAffinity PreFilter: conflicting metadata.name terms */
  bool has_names = false;       /* PreFilterResult.NodeNames */
  std::vector<int> names;
  bool ports_skip = true, pts_skip = true, ipa_skip = true;
  /* PodTopologySpread preFilterState (filtering.go:40-52) */
  std::vector<PtsCon> cons;
  std::vector<std::unordered_map<int, int>> tp_count; /* TpValueToMatchNum */
  std::vector<int> tp_min;                            /* CriticalPaths[i][0].MatchNum */
  /* InterPodAffinity preFilterState (filtering.go:44-56) */
  std::map<std::pair<int, int>, i64> existing_anti, affinity, anti;
};

struct Snapshot {
  Ctx c;
  std::vector<NodeState> nodes; /* ListNodeInfos() order: insertion order */
  int last_index = 0;           /* SchedulerPluginRunner.lastIndex (plugin_runner.go:34) */
  i64 filter_evals = 0;         /* RunFilterPlugins calls, for evals/s accounting */
  /* Fork/Revert (store/delta.go:561-588): journal of pre-fork nodes touched + size */
  bool forked = false;
  size_t fork_size = 0;
  std::vector<std::pair<int, NodeState>> undo;
  int fork_last_index = 0;

  explicit Snapshot(const cae_objects* o) : c(o) {}

  NodeState makeNode(int idx) const { /* NodeInfo for node-table row idx with its pods */
    const cae_objects* o = c.o;
    NodeState n;
    n.name = o->node_name[idx];
    n.labelset = o->node_labelset[idx];
    n.unschedulable = o->node_unschedulable[idx];
    n.taint_list = o->node_taint_list[idx];
    for (int r = 0; r < R; ++r) n.alloc[r] = o->node_alloc[(size_t)idx * R + r];
    n.allowed_pods = o->node_allowed_pods[idx];
    for (int i = o->node_pod_off[idx]; i < o->node_pod_off[idx + 1]; ++i) addPodRaw(n, o->node_pod_spec[i]);
    return n;
  }
  /* NodeInfo.AddPodInfo + update (types.go:347-357, 423-442) */
  void addPodRaw(NodeState& n, int spec) const {
    const cae_objects* o = c.o;
    n.pods.push_back(spec);
    for (int r = 0; r < R; ++r) n.requested[r] += o->ps_req[(size_t)spec * R + r];
    int pl = o->ps_port_list[spec];
    for (int i = o->port_off[pl]; i < o->port_off[pl + 1]; ++i)
      n.used_ports.push_back({o->port_ip[i], o->port_proto[i], o->port_num[i]});
  }
  void loadCluster() {
    for (int i = 0; i < c.o->num_cluster_nodes; ++i) nodes.push_back(makeNode(i));
  }
  void fork() { forked = true; fork_size = nodes.size(); undo.clear(); fork_last_index = last_index; }
  void revert(bool keep_last_index) {
    for (auto it = undo.rbegin(); it != undo.rend(); ++it) nodes[it->first] = it->second;
    undo.clear();
    nodes.resize(fork_size);
    forked = false;
    if (!keep_last_index) last_index = fork_last_index;
  }
  void touch(int idx) {
    if (forked && (size_t)idx < fork_size) {
      for (auto& u : undo) if (u.first == idx) return;
      undo.emplace_back(idx, nodes[idx]);
    }
  }
  void forceAddPod(int spec, int idx) { touch(idx); addPodRaw(nodes[idx], spec); }

  /* ---- PreFilter ------------------------------------------------------------------------ */
  void preFilter(int spec, PreFilter& pf) {
    const cae_objects* o = c.o;
    /* NodeAffinity.PreFilter (nodeaffinity/node_affinity.go:159-209) */
    int a = o->ps_naff[spec];
    if (a >= 0 && o->naff_has_required[a] && o->naff_term_off[a + 1] > o->naff_term_off[a]) {
      bool all_nodes = false, any = false;
      std::set<int> names;
      for (int t = o->naff_term_off[a]; t < o->naff_term_off[a + 1] && !all_nodes; ++t) {
        bool term_has = false, term_empty = false;
        int term_name = -1;
        for (int f = o->term_field_off[t]; f < o->term_field_off[t + 1]; ++f) {
          if (o->field_op[f] != CAE_OP_IN) continue;
          if (!term_has) { term_has = true; term_name = o->field_node_name[f]; }
          else if (term_name != o->field_node_name[f]) term_empty = true; /* intersection of singletons */
        }
        if (!term_has) { all_nodes = true; break; }
        any = true;
        if (!term_empty) names.insert(term_name);
      }
      if (!all_nodes && any) {
        if (names.empty()) { pf.failed = true; return; } /* errReasonConflict */
        pf.has_names = true;
        pf.names.assign(names.begin(), names.end());
      }
    }
    /* NodePorts.PreFilter: Skip if no host ports (nodeports/node_ports.go:75-84) */
    pf.ports_skip = (o->port_off[o->ps_port_list[spec] + 1] == o->port_off[o->ps_port_list[spec]]);
    preFilterPTS(spec, pf);
    preFilterIPA(spec, pf);
  }

  /* calPreFilterState (podtopologyspread/filtering.go:237-311) */
  void preFilterPTS(int spec, PreFilter& pf) {
    const cae_objects* o = c.o;
    int pl = o->ps_pts_list[spec];
    int nb = o->pts_off[pl], ne = o->pts_off[pl + 1];
    if (nb == ne) { pf.pts_skip = true; return; }
    pf.pts_skip = false;
    for (int i = nb; i < ne; ++i)
      pf.cons.push_back({o->pts_max_skew[i], o->pts_key[i], o->pts_selector[i], o->pts_min_domains[i],
                         o->pts_node_affinity_policy[i], o->pts_node_taints_policy[i]});
    size_t nc = pf.cons.size();
    pf.tp_count.assign(nc, {});
    int ns = o->ps_namespace[spec];
    for (const NodeState& n : nodes) {
      /* nodeLabelsMatchSpreadConstraints: ALL topology keys present (common.go:78-85) */
      bool all = true;
      std::vector<int> vals(nc);
      for (size_t i = 0; i < nc && all; ++i) all = c.nodeLabel(n, pf.cons[i].key, &vals[i]);
      if (!all) continue;
      for (size_t i = 0; i < nc; ++i) {
        const PtsCon& k = pf.cons[i];
        /* matchNodeInclusionPolicies (common.go:43-58) */
        if (k.aff_policy == CAE_POLICY_HONOR && !c.naffMatch(o->ps_naff[spec], n)) continue;
        if (k.taint_policy == CAE_POLICY_HONOR && c.hasUntoleratedTaint(n.taint_list, o->ps_tol_list[spec])) continue;
        /* countPodsMatchSelector (common.go:145-160) */
        int cnt = 0;
        if (!c.selEmpty(k.sel)) {
          for (int ps : n.pods) {
            if (o->ps_terminating[ps] || o->ps_namespace[ps] != ns) continue;
            if (c.selMatchesLabelset(k.sel, o->ps_labelset[ps])) ++cnt;
          }
        }
        pf.tp_count[i][vals[i]] += cnt;
      }
    }
    pf.tp_min.assign(nc, std::numeric_limits<int32_t>::max()); /* newCriticalPaths: MaxInt32 */
    for (size_t i = 0; i < nc; ++i)
      for (auto& kv : pf.tp_count[i]) pf.tp_min[i] = std::min(pf.tp_min[i], kv.second);
  }

  /* InterPodAffinity.PreFilter (interpodaffinity/filtering.go:274-309) */
  void preFilterIPA(int spec, PreFilter& pf) {
    const cae_objects* o = c.o;
    int al = o->ps_aff_list[spec], bl = o->ps_anti_list[spec];
    int ab = o->aff_off[al], ae = o->aff_off[al + 1], bb = o->aff_off[bl], be = o->aff_off[bl + 1];
    /* getExistingAntiAffinityCounts (:204-228) */
    for (const NodeState& n : nodes) {
      for (int ep : n.pods) {
        int el = o->ps_anti_list[ep];
        for (int t = o->aff_off[el]; t < o->aff_off[el + 1]; ++t) {
          if (!c.atermMatchesWithNsLabels(t, spec)) continue;
          int v;
          if (c.nodeLabel(n, o->aterm_key[t], &v)) pf.existing_anti[{o->aterm_key[t], v}] += 1;
        }
      }
    }
    /* getIncomingAffinityAntiAffinityCounts (:234-271) */
    if (ab != ae || bb != be) {
      for (const NodeState& n : nodes) {
        for (int ep : n.pods) {
          bool all = (ab != ae); /* podMatchesAllAffinityTerms: false for no terms (:187-199) */
          for (int t = ab; t < ae && all; ++t) all = c.incomingTermMatches(t, ep);
          if (all)
            for (int t = ab; t < ae; ++t) { int v; if (c.nodeLabel(n, o->aterm_key[t], &v)) pf.affinity[{o->aterm_key[t], v}] += 1; }
          for (int t = bb; t < be; ++t)
            if (c.incomingTermMatches(t, ep)) { int v; if (c.nodeLabel(n, o->aterm_key[t], &v)) pf.anti[{o->aterm_key[t], v}] += 1; }
        }
      }
    }
    pf.ipa_skip = pf.existing_anti.empty() && ab == ae && bb == be; /* :303-305 */
  }

  /* ---- Filter: default plugin order (K8S/apis/config/v1/default_plugins.go:34-52) ------------ */
  int runFilters(int spec, const PreFilter& pf, const NodeState& n) {
    const cae_objects* o = c.o;
    ++filter_evals;
    /* NodeUnschedulable (nodeunschedulable/node_unschedulable.go:142-160): the pod must tolerate
     * the taint {Key: node.kubernetes.io/unschedulable, Value: "", Effect: NoSchedule}. */
    if (n.unschedulable) {
      bool tol = false;
      int tl = o->ps_tol_list[spec];
      for (int i = o->tol_off[tl]; i < o->tol_off[tl + 1] && !tol; ++i) {
        if (o->tol_effect[i] != CAE_EFFECT_NONE && o->tol_effect[i] != CAE_EFFECT_NO_SCHEDULE) continue;
        /* taint {Key: node.kubernetes.io/unschedulable, Value: "", Effect: NoSchedule} */
        if (o->tol_key[i] >= 0 && o->tol_key[i] != o->unschedulable_taint_key) continue;
        if (o->tol_op[i] == CAE_TOL_EXISTS) tol = true;
        else if (o->tol_op[i] == CAE_TOL_EQUAL) tol = (o->tol_val[i] == -1);
      }
      if (!tol) return CAE_R_NODE_UNSCHEDULABLE;
    }
    /* NodeName (nodename/node_name.go:79-90) */
    if (o->ps_node_name[spec] >= 0 && o->ps_node_name[spec] != n.name) return CAE_R_NODE_NAME;
    /* TaintToleration (tainttoleration/taint_toleration.go:119-133) */
    if (c.hasUntoleratedTaint(n.taint_list, o->ps_tol_list[spec])) return CAE_R_TAINT;
    /* NodeAffinity.Filter (nodeaffinity/node_affinity.go:218-238); Skip if nothing to check (:166) */
    if (o->ps_naff[spec] >= 0 && !c.naffMatch(o->ps_naff[spec], n)) return CAE_R_NODE_AFFINITY;
    /* NodePorts.Filter (nodeports/node_ports.go:162-190) */
    if (!pf.ports_skip) {
      int pl = o->ps_port_list[spec];
      for (int i = o->port_off[pl]; i < o->port_off[pl + 1]; ++i)
        if (portConflict(n.used_ports, {o->port_ip[i], o->port_proto[i], o->port_num[i]})) return CAE_R_NODE_PORTS;
    }
    /* NodeResourcesFit fitsRequest (noderesources/fit.go:649-736) */
    {
      bool bad = ((int)n.pods.size() + 1 > n.allowed_pods);
      const i64* req = &o->ps_req[(size_t)spec * R];
      for (int r = 0; r < R && !bad; ++r)
        if (req[r] > 0 && req[r] > n.alloc[r] - n.requested[r]) bad = true;
      if (bad) return CAE_R_FIT;
    }
    /* PodTopologySpread.Filter (podtopologyspread/filtering.go:314-359) */
    if (!pf.pts_skip) {
      for (size_t i = 0; i < pf.cons.size(); ++i) {
        const PtsCon& k = pf.cons[i];
        int tpval;
        if (!c.nodeLabel(n, k.key, &tpval)) return CAE_R_PTS_MISSING_LABEL;
        int minm = pf.tp_min[i];
        if ((int)pf.tp_count[i].size() < k.min_domains) minm = 0; /* minMatchNum (:55-68) */
        int self = c.selMatchesLabelset(k.sel, o->ps_labelset[spec]) ? 1 : 0;
        auto it = pf.tp_count[i].find(tpval);
        int match = it == pf.tp_count[i].end() ? 0 : it->second;
        if ((i64)match + self - minm > k.max_skew) return CAE_R_PTS_SKEW;
      }
    }
    /* InterPodAffinity.Filter (interpodaffinity/filtering.go:412-432) */
    if (!pf.ipa_skip) {
      int al = o->ps_aff_list[spec], bl = o->ps_anti_list[spec];
      int ab = o->aff_off[al], ae = o->aff_off[al + 1], bb = o->aff_off[bl], be = o->aff_off[bl + 1];
      /* satisfyPodAffinity (:382-408) */
      bool pods_exist = true, ok = true;
      for (int t = ab; t < ae; ++t) {
        int v;
        if (c.nodeLabel(n, o->aterm_key[t], &v)) {
          auto it = pf.affinity.find({o->aterm_key[t], v});
          if (it == pf.affinity.end() || it->second <= 0) pods_exist = false;
        } else { ok = false; break; }
      }
      if (ok && !pods_exist) {
        bool self_all = (ab != ae);
        for (int t = ab; t < ae && self_all; ++t) self_all = c.incomingTermMatches(t, spec);
        ok = pf.affinity.empty() && self_all;
      }
      if (!ok) return CAE_R_IPA_AFFINITY;
      /* satisfyPodAntiAffinity (:367-379) */
      if (!pf.anti.empty())
        for (int t = bb; t < be; ++t) {
          int v;
          if (c.nodeLabel(n, o->aterm_key[t], &v)) {
            auto it = pf.anti.find({o->aterm_key[t], v});
            if (it != pf.anti.end() && it->second > 0) return CAE_R_IPA_ANTI_AFFINITY;
          }
        }
      /* satisfyExistingPodsAntiAffinity (:352-364): every label pair of the node */
      if (!pf.existing_anti.empty()) {
        for (auto& kv : pf.existing_anti) {
          if (kv.second <= 0) continue;
          int v;
          if (c.nodeLabel(n, kv.first.first, &v) && v == kv.first.second) return CAE_R_IPA_EXISTING_ANTI_AFFINITY;
        }
      }
    }
    return CAE_R_OK;
  }

  /* RunFiltersOnNode (predicate/plugin_runner.go:131-166) */
  int runFiltersOnNode(int spec, int idx) {
    PreFilter pf;
    preFilter(spec, pf);
    if (pf.failed) return CAE_R_PREFILTER_NODEAFFINITY;
    const NodeState& n = nodes[idx];
    if (pf.has_names && std::find(pf.names.begin(), pf.names.end(), n.name) == pf.names.end())
      return CAE_R_PREFILTER_NODEAFFINITY;
    return runFilters(spec, pf, n);
  }
  /* SchedulePod (predicate/predicate_snapshot.go:169-189) */
  int schedulePod(int spec, int idx) {
    int r = runFiltersOnNode(spec, idx);
    if (r == CAE_R_OK) forceAddPod(spec, idx);
    return r;
  }
  /* RunFiltersUntilPassingNode (plugin_runner.go:50-128), parallelism 1 */
  template <class Match>
  int schedulePodOnAnyNodeMatching(int spec, Match&& nodeMatches) {
    PreFilter pf;
    preFilter(spec, pf);
    if (pf.failed) return -1;
    int len = (int)nodes.size();
    if (len == 0) return -1;
    for (int i = 0; i < len; ++i) {
      int idx = (last_index + i) % len;
      const NodeState& n = nodes[idx];
      if (pf.has_names && std::find(pf.names.begin(), pf.names.end(), n.name) == pf.names.end()) continue;
      if (n.unschedulable) continue; /* :92-94 */
      if (!nodeMatches(n, idx)) continue;
      if (runFilters(spec, pf, n) == CAE_R_OK) {
        last_index = (idx + 1) % len;
        forceAddPod(spec, idx);
        return idx;
      }
    }
    return -1;
  }
};

/* calculatePodScore (estimator/decreasing_pod_orderer.go:61-88) */
double podScore(const cae_objects* o, int spec, int tmpl_node) {
  double score = 0;
  i64 cpu = o->ps_req[(size_t)spec * R + CAE_RES_CPU], mem = o->ps_req[(size_t)spec * R + CAE_RES_MEM];
  i64 acpu = o->node_alloc[(size_t)tmpl_node * R + CAE_RES_CPU], amem = o->node_alloc[(size_t)tmpl_node * R + CAE_RES_MEM];
  if (o->node_has_alloc_cpu[tmpl_node] && acpu > 0) score += (double)cpu / (double)acpu;
  if (o->node_has_alloc_mem[tmpl_node] && amem > 0) score += (double)mem / (double)amem;
  return score;
}

/* thresholdBasedEstimationLimiter, node-count part (estimator/threshold_based_limiter.go:26-69) */
struct Limiter {
  int max_nodes, nodes = 0;
  bool permissionToAddNode() {
    if (max_nodes < 0 || (max_nodes > 0 && nodes >= max_nodes)) return false;
    ++nodes;
    return true;
  }
};

struct EstimateResult {
  int node_count = 0, pod_count = 0;
  std::vector<int> order;       /* group ids in processing order */
  std::vector<int> sched;       /* per group id: pods scheduled */
  std::vector<int> placements;  /* node list index per scheduled pod, in scheduling order */
  int last_index = 0;           /* SchedulerPluginRunner.lastIndex when Estimate returns (raw: it survives the Revert) */
};

/* BinpackingNodeEstimator.Estimate (estimator/binpacking_estimator.go:97-139) */
void estimate(Snapshot& s, int tmpl, const std::vector<int>& groups_in, int max_nodes, int num_groups, EstimateResult& res,
              int last_index_in = 0) {
  const cae_objects* o = s.c.o;
  int tnode = o->num_cluster_nodes + tmpl;
  res.sched.assign(num_groups, 0);
  /* DecreasingPodOrderer.Order (decreasing_pod_orderer.go:46-58); stable for ties (see header) */
  std::vector<std::pair<double, int>> scored;
  for (int g : groups_in) {
    double sc = 0;
    if (o->group_off[g + 1] > o->group_off[g]) sc = podScore(o, o->pend_spec[o->group_off[g]], tnode);
    scored.emplace_back(sc, g);
  }
  std::stable_sort(scored.begin(), scored.end(), [](auto& a, auto& b) { return a.first > b.first; });
  Limiter lim{max_nodes};
  s.fork();
  s.last_index = last_index_in; /* 0 = fresh runner per Estimate (header note); a caller may carry the runner's value over
                                   (plugin_runner.go:81,123: used modulo the CURRENT list length until a scan places a pod) */
  int new_node_index = 0, last_node = -1; /* estimationState (:46-53) */
  int nodes_with_pods = 0;
  bool new_nodes_available = true;
  auto track = [&](int idx) { /* trackScheduledPod (:55-58) */
    if (!s.nodes[idx].has_sched) { s.touch(idx); s.nodes[idx].has_sched = true; ++nodes_with_pods; }
    res.placements.push_back(idx);
    ++res.pod_count;
  };
  for (auto& sg : scored) {
    int g = sg.second;
    res.order.push_back(g);
    int pb = o->group_off[g], pe = o->group_off[g + 1];
    /* tryToScheduleOnExistingNodes (:141-164) */
    int index = pb;
    for (; index < pe; ++index) {
      int spec = o->pend_spec[index];
      int idx = s.schedulePodOnAnyNodeMatching(spec, [&](const NodeState& n, int) { return n.is_new; });
      if (idx < 0) break;
      track(idx);
      ++res.sched[g];
    }
    if (!new_nodes_available) continue;
    /* tryToScheduleOnNewNodes (:168-247) */
    for (; index < pe; ++index) {
      int spec = o->pend_spec[index];
      bool found = false;
      if (last_node >= 0) {
        int r = s.schedulePod(spec, last_node);
        if (r == CAE_R_OK) { found = true; track(last_node); ++res.sched[g]; }
        /* isPodUsingHostNameTopologyKey && hasTopologyConstraintError (:186, :269-292) */
        const bool host_pts = o->ps_hostname_spread[spec] != 0;
        if (host_pts && r == CAE_R_PTS_SKEW) {
          int ln = last_node;
          int idx = s.schedulePodOnAnyNodeMatching(spec, [&](const NodeState&, int i) { return i != ln; });
          if (idx >= 0) { found = true; track(idx); ++res.sched[g]; }
        }
      }
      if (!found) {
        if (last_node >= 0 && !s.nodes[last_node].has_sched) break; /* :212 return true */
        if (!lim.permissionToAddNode()) { new_nodes_available = false; break; } /* :222 return false */
        /* addNewNodeToSnapshot (:249-265) + SanitizedNodeInfo (simulator/node_info_utils.go:90-139) */
        NodeState n = s.makeNode(tnode);
        n.sanitized = true;
        n.is_new = true;
        n.name = -2 - new_node_index;
        n.hostname_val = -2 - new_node_index;
        s.nodes.push_back(std::move(n));
        ++new_node_index;
        last_node = (int)s.nodes.size() - 1;
        int r = s.schedulePod(spec, last_node);
        if (r != CAE_R_OK) break; /* :238-240 */
        track(last_node);
        ++res.sched[g];
      }
    }
  }
  res.node_count = nodes_with_pods;
  res.last_index = s.last_index;
  s.revert(false);
}

/* SchedulablePodGroups (core/scaleup/orchestrator/orchestrator.go:603-638) for one (spec, template) */
int checkOnTemplate(Snapshot& s, int spec, int tmpl) {
  /* caller has forked and added the template node as the last node */
  (void)tmpl;
  return s.runFiltersOnNode(spec, (int)s.nodes.size() - 1);
}

}  // namespace

extern "C" {

const char* cao_version(void) { return "ca-oracle/1 (port of CA 1.35.0-rc.0 estimator+predicates)"; }

/* reasons[t * ncols + j]: j-th pod (dense) or j-th group exemplar vs template t.
 * mode 0 = dense over pending pods [p_begin, p_end); mode 1 = group exemplars (p_* ignored). */
int cao_feasibility(const cae_objects* o, int mode, int p_begin, int p_end, int t_begin, int t_end, uint8_t* reasons, int64_t* evals) {
  Snapshot s(o);
  s.loadCluster();
  int ncols = mode == 0 ? (p_end - p_begin) : o->num_groups;
  for (int t = t_begin; t < t_end; ++t) {
    s.fork();
    s.nodes.push_back(s.makeNode(o->num_cluster_nodes + t));
    for (int j = 0; j < ncols; ++j) {
      int spec;
      if (mode == 0) spec = o->pend_spec[p_begin + j];
      else {
        if (o->group_off[j + 1] == o->group_off[j]) { reasons[(size_t)(t - t_begin) * ncols + j] = CAE_R_OK; continue; }
        spec = o->pend_spec[o->group_off[j]];
      }
      reasons[(size_t)(t - t_begin) * ncols + j] = (uint8_t)checkOnTemplate(s, spec, t);
    }
    s.revert(true);
  }
  if (evals) *evals = s.filter_evals;
  return 0;
}

/* One Estimate call on a fresh fork of the cluster snapshot.
 * groups: caller's group order (what SchedulablePodGroups returned); outputs sized by num_groups. */
int cao_estimate(const cae_objects* o, int tmpl, const int32_t* groups, int n_groups, int max_nodes,
                 int32_t* node_count, int32_t* pod_count, int32_t* sched_count, int32_t* order,
                 int32_t* placements, int64_t* evals) {
  Snapshot s(o);
  s.loadCluster();
  EstimateResult res;
  estimate(s, tmpl, std::vector<int>(groups, groups + n_groups), max_nodes, o->num_groups, res);
  *node_count = res.node_count;
  *pod_count = res.pod_count;
  if (sched_count) std::copy(res.sched.begin(), res.sched.end(), sched_count);
  if (order) { std::fill(order, order + o->num_groups, -1); std::copy(res.order.begin(), res.order.end(), order); }
  if (placements) std::copy(res.placements.begin(), res.placements.end(), placements);
  if (evals) *evals = s.filter_evals;
  return 0;
}

/* The per-tick loop of ScaleUp for templates [t_begin, t_end): SchedulablePodGroups then Estimate
 * (orchestrator.go:145-150).  Outputs are [t_end - t_begin] / [(t_end - t_begin) * E]. */
int cao_estimate_all(const cae_objects* o, const int32_t* max_nodes, int t_begin, int t_end,
                     int32_t* node_count, int32_t* pod_count, int32_t* sched_count, int32_t* order, int64_t* evals) {
  Snapshot s(o);
  s.loadCluster();
  int E = o->num_groups;
  i64 ev = 0;
  for (int t = t_begin; t < t_end; ++t) {
    std::vector<int> feasible;
    s.fork();
    s.nodes.push_back(s.makeNode(o->num_cluster_nodes + t));
    for (int g = 0; g < E; ++g) {
      if (o->group_off[g + 1] == o->group_off[g]) continue;
      if (checkOnTemplate(s, o->pend_spec[o->group_off[g]], t) == CAE_R_OK) feasible.push_back(g);
    }
    s.revert(true);
    EstimateResult res;
    estimate(s, t, feasible, max_nodes ? max_nodes[t] : 0, E, res);
    size_t row = (size_t)(t - t_begin);
    node_count[row] = res.node_count;
    pod_count[row] = res.pod_count;
    if (sched_count) std::copy(res.sched.begin(), res.sched.end(), sched_count + row * E);
    if (order) { std::fill(order + row * E, order + (row + 1) * E, -1); std::copy(res.order.begin(), res.order.end(), order + row * E); }
  }
  ev = s.filter_evals;
  if (evals) *evals = ev;
  return 0;
}

/* cao_estimate_all with the plugin runner's lastIndex carried in (and out) per template; chain != 0: template t starts from the
 * value template t-1 ended with (one long-lived runner), last_index_in[t_begin] seeds the first. */
int cao_estimate_all_li(const cae_objects* o, const int32_t* max_nodes, const int32_t* last_index_in, int chain, int t_begin, int t_end,
                        int32_t* node_count, int32_t* pod_count, int32_t* sched_count, int32_t* order, int32_t* last_index_out) {
  Snapshot s(o);
  s.loadCluster();
  int E = o->num_groups;
  int carry = last_index_in ? last_index_in[t_begin] : 0;
  for (int t = t_begin; t < t_end; ++t) {
    std::vector<int> feasible;
    s.fork();
    s.nodes.push_back(s.makeNode(o->num_cluster_nodes + t));
    for (int g = 0; g < E; ++g) {
      if (o->group_off[g + 1] == o->group_off[g]) continue;
      if (checkOnTemplate(s, o->pend_spec[o->group_off[g]], t) == CAE_R_OK) feasible.push_back(g);
    }
    s.revert(true);
    EstimateResult res;
    const int li = chain ? carry : (last_index_in ? last_index_in[t] : 0);
    estimate(s, t, feasible, max_nodes ? max_nodes[t] : 0, E, res, li);
    carry = res.last_index;
    size_t row = (size_t)(t - t_begin);
    node_count[row] = res.node_count;
    pod_count[row] = res.pod_count;
    if (sched_count) std::copy(res.sched.begin(), res.sched.end(), sched_count + row * E);
    if (order) { std::fill(order + row * E, order + (row + 1) * E, -1); std::copy(res.order.begin(), res.order.end(), order + row * E); }
    if (last_index_out) last_index_out[row] = res.last_index;
  }
  return 0;
}

double cao_pod_score(const cae_objects* o, int spec, int tmpl) { return podScore(o, spec, o->num_cluster_nodes + tmpl); }

/* filterOutSchedulablePodListProcessor.filterOutSchedulableByPacking -> HintingSimulator.TrySchedulePods
 * (core/podlistprocessor/filter_out_schedulable.go:96-126, simulator/scheduling/hinting_simulator.go:53-135,
 * simulator/scheduling/similar_pods.go:59-112), breakOnFailure = false, on the cluster snapshot.
 *   pod_order   pending-pod indices in processing order (the caller's priority sort; Go's sort.Slice is unstable,
 *               so the order among equal priorities is the caller's to fix)
 *   hint_node   [num_pending] hinted cluster node of a pod (Hints.Get), -1 = none; NULL = no hints
 *   sim_class   [num_pending] id of (controller UID, labels, spec) for pods with a controller that is not a
 *               DaemonSet, -1 otherwise (ControllerRef nil / IsDaemonSetPod); NULL = none
 *   class_ctrl  [classes] controller id of a class
 *   node_ok     [N] isNodeAcceptable, NULL = ScheduleAnywhere
 * Outputs: assigned[num_pending] = cluster node the pod was placed on, -1 = stays unschedulable;
 * the runner's lastIndex afterwards; OverflowingControllerCount. */
int cao_filter_schedulable(const cae_objects* o, const int32_t* pod_order, int n_pods, const int32_t* hint_node,
                           const int32_t* sim_class, const int32_t* class_ctrl, const uint8_t* node_ok, int last_index_in,
                           int break_on_failure, int32_t* assigned, int32_t* last_index_out, int32_t* overflowing) {
  Snapshot s(o);
  s.loadCluster();
  s.last_index = last_index_in;
  const int N = o->num_cluster_nodes;
  for (int i = 0; i < o->num_pending; ++i) assigned[i] = -1;
  std::map<int, std::vector<int>> items;   /* controller -> classes known unschedulable (SimilarPodsScheduling.items) */
  std::set<int> over;                      /* overflowingControllers */
  for (int k = 0; k < n_pods; ++k) {
    const int pod = pod_order[k];
    const int spec = o->pend_spec[pod];
    int where = -1;
    /* tryScheduleUsingHints (:80-106) */
    const int h = hint_node ? hint_node[pod] : -1;
    if (h >= 0 && h < N && (!node_ok || node_ok[h]) && s.schedulePod(spec, h) == CAE_R_OK) where = h;
    if (where < 0) {
      /* trySchedule (:110-130) */
      const int cls = sim_class ? sim_class[pod] : -1;
      bool similar_unsched = false;
      if (cls >= 0) {
        auto it = items.find(class_ctrl[cls]);
        if (it != items.end()) similar_unsched = std::find(it->second.begin(), it->second.end(), cls) != it->second.end();
      }
      if (!similar_unsched) {
        where = s.schedulePodOnAnyNodeMatching(spec, [&](const NodeState&, int idx) { return !node_ok || node_ok[idx]; });
        if (where < 0 && cls >= 0) { /* SetUnschedulable (similar_pods.go:87-104) */
          std::vector<int>& pm = items[class_ctrl[cls]];
          if ((int)pm.size() >= 10) over.insert(class_ctrl[cls]);
          else pm.push_back(cls);
        }
      }
    }
    assigned[pod] = where;
    if (where < 0 && break_on_failure) break; /* hinting_simulator.go:71-73 */
  }
  if (last_index_out) *last_index_out = s.last_index;
  if (overflowing) *overflowing = (int32_t)over.size();
  return 0;
}

/* getMinLimit (estimator/threshold_based_limiter.go:45-53) */
int64_t cao_get_min_limit(int64_t base, int64_t target) {
  if (base < 0 || target < 0) return -1;
  if ((base == 0 || base > target) && target > 0) return target;
  return base;
}
/* clusterCapacityThreshold.NodeLimit (estimator/cluster_capacity_threshold.go:33-41); has_ctx=0 -> nil context */
int cao_cluster_capacity_limit(int has_ctx, int cluster_max_node_limit, int current_node_count) {
  if (!has_ctx || cluster_max_node_limit == 0) return 0;
  if (cluster_max_node_limit < 0 || cluster_max_node_limit <= current_node_count) return -1;
  return cluster_max_node_limit - current_node_count;
}
/* sngCapacityThreshold.NodeLimit (estimator/sng_capacity_threshold.go:34-60): groups = this + similar */
int cao_sng_capacity_limit(int has_ctx, const int32_t* max_size, const int32_t* target_size, int n) {
  if (!has_ctx) return 0;
  int total = 0;
  for (int i = 0; i < n; ++i) { int c = max_size[i] - target_size[i]; if (c > 0) total += c; }
  return total <= 0 ? -1 : total;
}
/* StartEstimation + repeated PermissionToAddNode (threshold_based_limiter.go:34-69), no duration:
 * returns how many of `asks` consecutive requests are granted. */
int cao_limiter_grants(const int32_t* node_limits, int n_thresholds, int asks) {
  i64 mx = 0;
  for (int i = 0; i < n_thresholds; ++i) mx = cao_get_min_limit(mx, node_limits[i]);
  Limiter lim{(int)mx};
  int granted = 0;
  for (int i = 0; i < asks; ++i) if (lim.permissionToAddNode()) ++granted; else break;
  return granted;
}

/* Expander filters over options (one per template with node_count>0; pods = scheduled prefix per group).
 * least-waste (expander/waste/waste.go:37-73), most-pods (mostpods/mostpods.go:33-54),
 * least-nodes (leastnodes/leastnodes.go:35-61), chain (factory/chain.go:36-45, without the random fallback).
 * in/out mask[T]: 1 = option still in the set. */
static void sumRequests(const cae_objects* o, const int32_t* sched_row, i64* cpu, i64* mem) {
  *cpu = 0; *mem = 0;
  for (int g = 0; g < o->num_groups; ++g) {
    if (sched_row[g] <= 0) continue;
    /* prefix of the group; pods of a group may differ, so sum the actual prefix */
    for (int p = o->group_off[g]; p < o->group_off[g] + sched_row[g]; ++p) {
      int spec = o->pend_spec[p];
      *cpu += o->ps_req[(size_t)spec * R + CAE_RES_CPU];
      *mem += o->ps_req[(size_t)spec * R + CAE_RES_MEM];
    }
  }
}
double cao_waste_score(const cae_objects* o, int t, int node_count, const int32_t* sched_row) {
  i64 rcpu, rmem;
  sumRequests(o, sched_row, &rcpu, &rmem);
  int n = o->num_cluster_nodes + t;
  i64 acpu = o->node_cap_cpu[n] * (i64)node_count, amem = o->node_cap_mem[n] * (i64)node_count;
  double wcpu = (double)(acpu - rcpu) / (double)acpu;
  double wmem = (double)(amem - rmem) / (double)amem;
  return wcpu + wmem;
}
/* ---- price expander (expander/price/price.go:90-183) ------------------------------------------------------------
 * math.Tanh / math.Exp restated from Go's pure-Go implementations (src/math/tanh.go, src/math/exp.go); this file is
 * compiled without FMA contraction (x86-64 baseline), like Go's amd64 code generator at GOAMD64=v1. */
static double goExp(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10, Log2e = 1.44269504088896338700e+00;
  const double Overflow = 7.09782712893383973096e+02, Underflow = -7.45133219101941108420e+02, NearZero = 1.0 / (1 << 28);
  if (std::isnan(x) || (std::isinf(x) && x > 0)) return x;
  if (std::isinf(x)) return 0;
  if (x > Overflow) return std::numeric_limits<double>::infinity();
  if (x < Underflow) return 0;
  if (-NearZero < x && x < NearZero) return 1 + x;
  int k = 0;
  if (x < 0) k = (int)(Log2e * x - 0.5);
  else if (x > 0) k = (int)(Log2e * x + 0.5);
  double hi = x - (double)k * Ln2Hi;
  double lo = (double)k * Ln2Lo;
  const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  double r = hi - lo;
  double t = r * r;
  double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
  return std::ldexp(y, k);
}
static double goTanh(double x) {
  const double MAXLOG = 8.8029691931113054295988e+01;
  double z = std::fabs(x);
  if (z > 0.5 * MAXLOG) return x < 0 ? -1 : 1;
  if (z >= 0.625) {
    double s = goExp(2 * z);
    z = 1 - 2 / (s + 1);
    return x < 0 ? -z : z;
  }
  if (x == 0) return x;
  const double P[3] = {-9.64399179425052238628e-1, -9.92877231001918586564e1, -1.61468768441708447952e3};
  const double Q[3] = {1.12811678491632931402e2, 2.23548839060100448583e3, 4.84406305325125486048e3};
  double s = x * x;
  return x + x * s * ((P[0] * s + P[1]) * s + P[2]) / (((s + Q[0]) * s + Q[1]) * s + Q[2]);
}
double cao_go_tanh(double x) { return goTanh(x); }
/* score of every option; sched/order rows as Estimate returned them (scheduledPods = groups in processing order, a prefix each) */
int cao_price_scores(const cae_objects* o, const double* node_price, const double* pod_price, const double* unfitness,
                     const uint8_t* has_gpu, const uint8_t* exists, double stabilization, int64_t preferred_cpu_milli,
                     const int32_t* node_count, const int32_t* sched, const int32_t* order, double* score) {
  int T = o->num_templates, E = o->num_groups, N = o->num_cluster_nodes;
  for (int t = 0; t < T; ++t) {
    score[t] = 0;
    if (node_count[t] <= 0) continue;
    double totalNodePrice = node_price[t] * (double)node_count[t];
    double totalPodPrice = 0;
    for (int gi = 0; gi < E; ++gi) {
      int g = order[(size_t)t * E + gi];
      if (g < 0) break;
      for (int p = o->group_off[g]; p < o->group_off[g] + sched[(size_t)t * E + g]; ++p) totalPodPrice += pod_price[o->pend_spec[p]];
    }
    double priceSubScore = (totalNodePrice + stabilization) / (totalPodPrice + stabilization);
    double nodeUnfitness;
    if (unfitness) nodeUnfitness = unfitness[t];
    else { /* SimpleNodeUnfitness (preferred.go:87-92) */
      double pref = (double)preferred_cpu_milli, ev = (double)o->node_cap_cpu[N + t];
      double a = pref / ev, b = ev / pref;
      nodeUnfitness = (std::isnan(a) || std::isnan(b)) ? a + b : std::max(a, b);
    }
    double supressedUnfitness = (nodeUnfitness - 1.0) * (1.0 - goTanh((double)(node_count[t] - 1) / 15.0)) + 1.0;
    if (has_gpu && has_gpu[t]) supressedUnfitness = 1000.0; /* gpuUnfitnessOverride */
    double optionScore = supressedUnfitness * priceSubScore;
    if (exists && !exists[t]) optionScore *= 2.0; /* notExistCoeficient */
    score[t] = optionScore;
  }
  return 0;
}
/* the chain with price (price.go:166-173) and priority (priority.go:119-165) filters over given vectors */
int cao_expander_ex(int T, const int32_t* chain, int chain_len, const int32_t* node_count, const int32_t* pod_count,
                    const double* waste, const double* price, const uint8_t* price_error, const int32_t* priority, uint8_t* mask) {
  std::vector<int> opts;
  for (int t = 0; t < T; ++t) if (node_count[t] > 0) opts.push_back(t);
  for (int c = 0; c < chain_len; ++c) {
    std::vector<int> best;
    if (chain[c] == CAE_EXP_LEAST_WASTE) {
      double least = 0;
      for (int t : opts) {
        if (waste[t] == least) best.push_back(t);
        if (best.empty() || waste[t] < least) { least = waste[t]; best.assign(1, t); }
      }
    } else if (chain[c] == CAE_EXP_MOST_PODS) {
      int mx = 0;
      for (int t : opts) {
        if (pod_count[t] == mx) { best.push_back(t); continue; }
        if (pod_count[t] > mx) { mx = pod_count[t]; best.assign(1, t); }
      }
    } else if (chain[c] == CAE_EXP_LEAST_NODES) {
      int least = std::numeric_limits<int>::max();
      for (int t : opts) {
        if (node_count[t] == 0) continue;
        if (node_count[t] == least) { best.push_back(t); continue; }
        if (node_count[t] < least) { least = node_count[t]; best.assign(1, t); }
      }
    } else if (chain[c] == CAE_EXP_PRICE) {
      double bestOptionScore = 0.0;
      for (int t : opts) {
        if (price_error && price_error[t]) continue; /* continue nextoption */
        if (best.empty() || bestOptionScore == price[t]) { best.push_back(t); bestOptionScore = price[t]; }
        else if (bestOptionScore > price[t]) { best.assign(1, t); bestOptionScore = price[t]; }
      }
    } else if (chain[c] == CAE_EXP_PRIORITY) {
      int maxPrio = -1;
      for (int t : opts) {
        if (priority[t] < 0) continue;          /* !found: "The group won't be used" */
        if (priority[t] < maxPrio) continue;
        if (priority[t] > maxPrio) { maxPrio = priority[t]; best.clear(); }
        best.push_back(t);
      }
      if (best.empty()) best = opts;            /* "No options filtered." */
    } else return 1;
    opts = best;
    if (opts.size() == 1) break;
  }
  for (int t = 0; t < T; ++t) mask[t] = 0;
  for (int t : opts) mask[t] = 1;
  return 0;
}

int cao_expander(const cae_objects* o, const int32_t* chain, int chain_len, const int32_t* node_count,
                 const int32_t* pod_count, const int32_t* sched_count, uint8_t* mask, double* waste_out) {
  int T = o->num_templates, E = o->num_groups;
  std::vector<int> opts;
  for (int t = 0; t < T; ++t) { mask[t] = node_count[t] > 0; if (mask[t]) opts.push_back(t); }
  if (waste_out) for (int t = 0; t < T; ++t) waste_out[t] = node_count[t] > 0 ? cao_waste_score(o, t, node_count[t], sched_count + (size_t)t * E) : 0.0;
  for (int c = 0; c < chain_len; ++c) {
    std::vector<int> best;
    if (chain[c] == CAE_EXP_LEAST_WASTE) {
      double least = 0; /* var leastWastedScore float64; leastWastedOptions == nil <=> best.empty() */
      for (int t : opts) {
        double w = cao_waste_score(o, t, node_count[t], sched_count + (size_t)t * E);
        if (w == least) best.push_back(t); /* equality is tested first (waste.go:58-60) */
        if (best.empty() || w < least) { least = w; best.assign(1, t); }
      }
    } else if (chain[c] == CAE_EXP_MOST_PODS) {
      int mx = 0;
      for (int t : opts) {
        if (pod_count[t] == mx) { best.push_back(t); continue; }
        if (pod_count[t] > mx) { mx = pod_count[t]; best.assign(1, t); }
      }
    } else if (chain[c] == CAE_EXP_LEAST_NODES) {
      int least = std::numeric_limits<int>::max();
      for (int t : opts) {
        if (node_count[t] == 0) continue;
        if (node_count[t] == least) { best.push_back(t); continue; }
        if (node_count[t] < least) { least = node_count[t]; best.assign(1, t); }
      }
    } else return 1;
    opts = best;
    if (opts.size() == 1) break; /* chain returns early on a single survivor */
  }
  for (int t = 0; t < T; ++t) mask[t] = 0;
  for (int t : opts) mask[t] = 1;
  return 0;
}

}  // extern "C"
