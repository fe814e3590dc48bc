// extender_json.h -- JSON wire format of the scheduler-extender verbs, for the hot path's callers
// (reference: pkg/routes/routes.go:39-163, k8s.io/kube-scheduler v0.18.0 extender/v1).
//
//   ExtenderArgs          {"pod": v1.Pod, "nodes": ..., "nodenames": [..]}            routes.go:46-64
//   ExtenderFilterResult  {"nodenames": [..], "failedNodes": {name: msg}, "error": s} routes.go:72-83
//   HostPriorityList      [{"host": name, "score": int64}, ...]                        routes.go:104-116
//   ExtenderBindingArgs   {"podName","podNamespace","podUID","node"}                   routes.go:131-143
//   ExtenderBindingResult {"error": s}                                                 routes.go:149-162
//
// Dependency-free C++17; encoders reproduce Go's encoding/json output byte for byte (omitempty,
// sorted map keys, HTML-safe escaping).  A filter request for 10^5 candidate nodes is ~2 MB of names:
// names are interned to dense int32 ids in one pass without per-name allocation.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "resource_scheduler.h"

namespace egs {

// name -> dense id (the ids libegs uses); open addressing over one string arena
class NodeInterner {
 public:
  NodeInterner();
  int Intern(std::string_view name);                 // existing or new id
  int Find(std::string_view name) const;             // -1 when unknown
  const std::string &Name(int id) const { return names_[id]; }
  int size() const { return (int)names_.size(); }

 private:
  void Grow();
  std::vector<std::string> names_;
  std::vector<int32_t> table_;                       // -1 empty
};

struct ExtenderArgs {
  Pod pod;
  bool has_nodenames = false;                        // routes.go:59-64 rejects a request without them
  std::vector<int32_t> node_ids;                     // interned, request order
};
struct BindingArgs { std::string pod_name, pod_namespace, pod_uid, node; };

// resource.Quantity.Value(): the value rounded UP to an integer ("100", "4", "1Gi", "1500m", "2e3", "1.5")
bool ParseQuantityValue(std::string_view q, int64_t *out);

// returns "" or a parse error; unknown JSON members are skipped
std::string ParseExtenderArgs(std::string_view json, NodeInterner *nodes, ExtenderArgs *out);
std::string ParseBindingArgs(std::string_view json, BindingArgs *out);

// has_node_names == false: NodeNames is a nil pointer (every error path) and the member is omitted
std::string EncodeFilterResult(const std::vector<std::string> &node_names, const std::map<std::string, std::string> &failed,
                               const std::string &error, bool has_node_names = true);
std::string EncodeHostPriorityList(const std::vector<std::pair<std::string, int64_t>> &scores);
std::string EncodeBindingResult(const std::string &error);
void AppendJsonString(std::string *out, std::string_view s);   // Go encoding/json string escaping

}  // namespace egs
