// host_capi.cc -- flat C entry points over egs::CudaUnitScheduler so that tests (ctypes) can drive
// the C++ mirror of the reference's plugin interface.  A fake apiserver = a table of registered nodes.
#include <cstring>
#include <map>
#include <string>

#include "resource_scheduler.h"

using namespace egs;

struct HostCtx {
  std::map<std::string, NodeInfo> cluster;     // the "apiserver"
  CudaUnitScheduler *sch = nullptr;
  std::string buf;
};

static const char *ret(HostCtx *c, const std::string &s) { c->buf = s; return c->buf.c_str(); }

extern "C" {

void *egsh_create(int policy, int max_nodes, int device) {
  HostCtx *c = new HostCtx();
  c->sch = new CudaUnitScheduler(policy, max_nodes, device, [c](const std::string &name, NodeInfo *out) -> std::string {
    auto it = c->cluster.find(name);
    if (it == c->cluster.end()) return "nodes \"" + name + "\" not found";
    *out = it->second;
    return "";
  });
  if (!c->sch->ok()) { delete c->sch; delete c; return nullptr; }
  return c;
}
void egsh_destroy(void *h) { HostCtx *c = (HostCtx *)h; delete c->sch; delete c; }
void egsh_register_node(void *h, const char *name, int64_t core_alloc, int64_t mem_alloc) {
  NodeInfo &n = ((HostCtx *)h)->cluster[name];
  n.core_allocatable = core_alloc; n.mem_allocatable = mem_alloc;
}
void egsh_register_assumed_pod(void *h, const char *node, void *pod) {
  ((HostCtx *)h)->cluster[node].assumed_pods.push_back(*(Pod *)pod);
}

void *egsh_pod_new(const char *ns, const char *name, const char *uid, const char *node_name) {
  Pod *p = new Pod(); p->ns = ns; p->name = name; p->uid = uid; p->node_name = node_name; return p;
}
void egsh_pod_free(void *p) { delete (Pod *)p; }
// has_* == 0 leaves the key out of Requests (absent keys read as 0, pod.go:94-108)
void egsh_pod_add_container(void *p, const char *cname, int has_core, int64_t core, int has_mem, int64_t mem) {
  Container c; c.name = cname;
  if (has_core) c.requests[kResourceGPUCore] = core;
  if (has_mem) c.requests[kResourceGPUMemory] = mem;
  ((Pod *)p)->containers.push_back(c);
}
void egsh_pod_set_annotation(void *p, const char *k, const char *v) { ((Pod *)p)->annotations[k] = v; }
// annotations and labels as "A\tkey\tvalue\n" / "L\tkey\tvalue\n"
const char *egsh_pod_meta(void *h, void *p) {
  std::string s;
  for (auto &kv : ((Pod *)p)->annotations) s += "A\t" + kv.first + "\t" + kv.second + "\n";
  for (auto &kv : ((Pod *)p)->labels) s += "L\t" + kv.first + "\t" + kv.second + "\n";
  return ret((HostCtx *)h, s);
}
int egsh_handles(void *p) { return CudaUnitScheduler::Handles(*(Pod *)p) ? 1 : 0; }

// "E\t<error>\n" or per node "F\t<name>\n" (filtered, input order) / "X\t<name>\t<message>\n" (failedNodes)
const char *egsh_assume(void *h, void *pod, const char **nodes, int n) {
  HostCtx *c = (HostCtx *)h;
  std::vector<std::string> names(nodes, nodes + n), filtered;
  std::map<std::string, std::string> failed;
  std::string e = c->sch->Assume(names, *(Pod *)pod, &filtered, &failed);
  if (!e.empty()) return ret(c, "E\t" + e + "\n");
  std::string s;
  for (auto &f : filtered) s += "F\t" + f + "\n";
  for (auto &kv : failed) s += "X\t" + kv.first + "\t" + kv.second + "\n";
  return ret(c, s);
}
void egsh_score(void *h, void *pod, const char **nodes, int n, int64_t *out) {
  std::vector<std::string> names(nodes, nodes + n);
  auto sc = ((HostCtx *)h)->sch->Score(names, *(Pod *)pod);
  for (int i = 0; i < n; i++) out[i] = sc[i];
}
const char *egsh_bind(void *h, const char *node, void *pod) { return ret((HostCtx *)h, ((HostCtx *)h)->sch->Bind(node, (Pod *)pod)); }
const char *egsh_add_pod(void *h, void *pod) { return ret((HostCtx *)h, ((HostCtx *)h)->sch->AddPod(*(Pod *)pod)); }
const char *egsh_forget_pod(void *h, void *pod) { return ret((HostCtx *)h, ((HostCtx *)h)->sch->ForgetPod(*(Pod *)pod)); }
int egsh_known_pod(void *h, void *pod) { return ((HostCtx *)h)->sch->KnownPod(*(Pod *)pod); }
int egsh_released_pod(void *h, void *pod) { return ((HostCtx *)h)->sch->ReleasedPod(*(Pod *)pod); }
const char *egsh_status(void *h) { return ret((HostCtx *)h, ((HostCtx *)h)->sch->Status()); }

}  // extern "C"

// ---- extender JSON codec (no GPU needed) ----------------------------------------------------------
#include "extender_json.h"

struct JsonCtx { NodeInterner nodes; ExtenderArgs args; BindingArgs bind; std::string buf; };

extern "C" {
void *egsj_new() { return new JsonCtx(); }
void egsj_free(void *c) { delete (JsonCtx *)c; }
// returns "" or the error; afterwards egsj_* accessors read the parsed request
const char *egsj_parse_args(void *c, const char *json, int64_t len) {
  JsonCtx *j = (JsonCtx *)c;
  j->buf = ParseExtenderArgs(std::string_view(json, (size_t)len), &j->nodes, &j->args);
  return j->buf.c_str();
}
int egsj_has_nodenames(void *c) { return ((JsonCtx *)c)->args.has_nodenames; }
int egsj_n_nodes(void *c) { return (int)((JsonCtx *)c)->args.node_ids.size(); }
const int32_t *egsj_node_ids(void *c) { return ((JsonCtx *)c)->args.node_ids.data(); }
int egsj_interned(void *c) { return ((JsonCtx *)c)->nodes.size(); }
const char *egsj_node_name(void *c, int id) { return ((JsonCtx *)c)->nodes.Name(id).c_str(); }
// "ns\tname\tuid\tnodeName\n" then per container "C\tname\thasCore\tcore\thasMem\tmem\n", annotations "A\tk\tv\n"
const char *egsj_pod_dump(void *c) {
  JsonCtx *j = (JsonCtx *)c;
  const Pod &p = j->args.pod;
  std::string s = p.ns + "\t" + p.name + "\t" + p.uid + "\t" + p.node_name + "\n";
  for (const auto &ct : p.containers) {
    auto ci = ct.requests.find(kResourceGPUCore), mi = ct.requests.find(kResourceGPUMemory);
    s += "C\t" + ct.name + "\t" + (ci != ct.requests.end() ? "1\t" + std::to_string(ci->second) : std::string("0\t0")) + "\t" +
         (mi != ct.requests.end() ? "1\t" + std::to_string(mi->second) : std::string("0\t0")) + "\n";
  }
  for (const auto &kv : p.annotations) s += "A\t" + kv.first + "\t" + kv.second + "\n";
  j->buf = s;
  return j->buf.c_str();
}
const char *egsj_parse_binding(void *c, const char *json, int64_t len) {
  JsonCtx *j = (JsonCtx *)c;
  std::string e = ParseBindingArgs(std::string_view(json, (size_t)len), &j->bind);
  j->buf = e.empty() ? ("\t" + j->bind.pod_name + "\t" + j->bind.pod_namespace + "\t" + j->bind.pod_uid + "\t" + j->bind.node) : e;
  return j->buf.c_str();
}
int egsj_quantity(const char *q, int64_t *out) { return ParseQuantityValue(q, out) ? 1 : 0; }
// names / failed as "\n"-separated lists ("name\tmsg" for failed)
const char *egsj_encode_filter(void *c, const char *names, const char *failed, const char *error) {
  JsonCtx *j = (JsonCtx *)c;
  std::vector<std::string> nn; std::map<std::string, std::string> ff;
  auto split = [](const std::string &s, auto f) { size_t b = 0; while (b < s.size()) { size_t e = s.find('\n', b); if (e == std::string::npos) e = s.size(); if (e > b) f(s.substr(b, e - b)); b = e + 1; } };
  split(names, [&](const std::string &l) { nn.push_back(l); });
  split(failed, [&](const std::string &l) { size_t t = l.find('\t'); ff[l.substr(0, t)] = t == std::string::npos ? "" : l.substr(t + 1); });
  j->buf = EncodeFilterResult(nn, ff, error);
  return j->buf.c_str();
}
const char *egsj_encode_priorities(void *c, const char *names, const int64_t *scores, int n) {
  JsonCtx *j = (JsonCtx *)c;
  std::vector<std::pair<std::string, int64_t>> v;
  std::string s(names); size_t b = 0;
  for (int i = 0; i < n; i++) { size_t e = s.find('\n', b); if (e == std::string::npos) e = s.size(); v.emplace_back(s.substr(b, e - b), scores[i]); b = e + 1; }
  j->buf = EncodeHostPriorityList(v);
  return j->buf.c_str();
}
const char *egsj_encode_binding(void *c, const char *error) { JsonCtx *j = (JsonCtx *)c; j->buf = EncodeBindingResult(error); return j->buf.c_str(); }
}
