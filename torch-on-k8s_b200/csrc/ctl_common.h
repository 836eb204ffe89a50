// Shared declarations of the C++ control plane (ctl_job.cpp, ctl_coord.cpp, ctl_elastic.cpp).
#pragma once
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "json.h"
#include "tok_internal.h"

// A parsed TorchJob: the manifest as a JSON document plus the fields the reference marks `json:"-"`
// (TaskSpec.DependsOn, apis/train/v1alpha1/torchjob_types.go:103).
struct tok_job {
  tok::json::Value doc;
  // task type -> [(upstream task type, phase)]
  std::map<std::string, std::vector<std::pair<std::string, std::string>>> depends;
  bool defaulted = false;
};

namespace tok {

bool gate(unsigned g);
std::string lower(const std::string& s);
bool equal_fold(const std::string& a, const std::string& b);
char* dup_cstr(const std::string& s);
int out_json(const json::Value& v, char** out);
std::string gen_general_name(const std::string& job, const std::string& task_type,
                             const std::string& index);
json::Value* task_specs(tok_job* j);
const json::Value* task_specs(const tok_job* j);
int64_t num_tasks(const json::Value& task_spec);
int64_t total_tasks_excluding_aimaster(const json::Value& specs);
int64_t replica_slots(const std::string& task_type, const json::Value& task_spec);
std::string job_name(const tok_job* j);
std::string job_namespace(const tok_job* j);
json::Value& status_of(tok_job* j);
bool has_condition(const json::Value& status, const std::string& type);
void set_condition(json::Value& status, const std::string& type, const std::string& reason,
                   const std::string& message, const std::string& now);

}  // namespace tok
