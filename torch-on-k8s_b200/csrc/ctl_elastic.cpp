// ctl_elastic.cpp — torchelastic replica-count policy (the "pkg/torchelastic rescaling" of
// north_star; the code lives in controllers/train/torchelastic/):
//   decision procedure     elastic_scale.go:42-246   (one pass per managed job, every 30 s there)
//   status writers         job.go:41-92
//   continue test          job.go:94-100  lat[last]/last > lat[cur]/cur on the metricCount-th samples
//   doubling               job.go:102-104 (clamped to numMaxReplicas here: the reference does not)
//   log-line contract      observation.go:40-85
// The policy output drives tok_comm_reform (in-place re-form) instead of container restarts.
#include <stdlib.h>
#include <string.h>

#include <map>
#include <regex>
#include <string>
#include <vector>

#include "ctl_common.h"

using namespace tok;
using json::Value;

struct tok_elastic {
  int metric_count = 5;  // elastictorchjob_controller.go:60
  // job key -> replicas -> observed latencies (r.metrics)
  std::map<std::string, std::map<int64_t, std::vector<double>>> metrics;
};

namespace {

Value* worker_status(tok_job* j) {
  Value* st = j->doc.find("status");
  Value* es = st ? st->find("elasticScalingStatues") : nullptr;  // wire name incl. typo
  return (es && es->is_object()) ? es->find("Worker") : nullptr;
}

int result(const char* action, int64_t replicas, const Value* ws, const std::string& msg, char** out) {
  Value r = Value::object();
  r["action"] = Value::str(action);
  r["replicas"] = Value::integer(replicas);
  r["condition"] = Value::str(ws && ws->find("elasticCondition") ? ws->find("elasticCondition")->as_string() : "");
  r["continue"] = Value::boolean(ws && ws->find("continue") ? ws->find("continue")->as_bool() : false);
  r["message"] = Value::str(msg);
  return out_json(r, out);
}

}  // namespace

extern "C" {

int tok_elastic_create(int metric_count, tok_elastic_t** out) {
  if (!out) return fail(TOK_ERR_INVALID, "elastic out pointer is null");
  tok_elastic* e = new tok_elastic();
  if (metric_count > 0) e->metric_count = metric_count;
  *out = e;
  return TOK_OK;
}

void tok_elastic_destroy(tok_elastic_t* e) { delete e; }

// observation.go:54-76.  Tab-separated progress line of the pytorch/examples imagenet trainer:
// field 0 holds "Epoch", the first [0-9]{1,2} is the epoch and the first [0-9]{2,4} the batch;
// field 1's first [0-9]{1,2}.[0-9]{3} is the batch latency in seconds; field 5's first
// [0-9]{1,2}.[0-9]{1,2} the accuracy.
int tok_elastic_parse_log(const char* line, char** out) {
  if (!line) return fail(TOK_ERR_INVALID, "log line is null");
  std::vector<std::string> f;
  {
    std::string cur;
    for (const char* p = line; *p; ++p) {
      if (*p == '\t') {
        f.push_back(cur);
        cur.clear();
      } else if (*p != '\n' && *p != '\r') {
        cur += *p;
      }
    }
    f.push_back(cur);
  }
  if (f[0].find("Epoch") == std::string::npos)
    return fail(TOK_ERR_INVALID, "current line of log is not a torchelastic training log");
  if (f.size() < 6)
    return fail(TOK_ERR_INVALID, "torchelastic log line has %zu tab-separated fields, need 6", f.size());
  static const std::regex epoch_re("[0-9]{1,2}"), batch_re("[0-9]{2,4}"),
      train_re("[0-9]{1,2}.[0-9]{3}"), acc_re("[0-9]{1,2}.[0-9]{1,2}");
  std::smatch m;
  Value r = Value::object();
  if (!std::regex_search(f[0], m, epoch_re)) return fail(TOK_ERR_INVALID, "no epoch in %s", f[0].c_str());
  r["epoch"] = Value::integer(atoi(m.str().c_str()));
  if (!std::regex_search(f[0], m, batch_re)) return fail(TOK_ERR_INVALID, "no batch in %s", f[0].c_str());
  r["batch"] = Value::integer(atoi(m.str().c_str()));
  if (!std::regex_search(f[1], m, train_re)) return fail(TOK_ERR_INVALID, "no latency in %s", f[1].c_str());
  const double lat = strtod(m.str().c_str(), nullptr);  // ParseFloat fails on a non-'.' separator -> 0
  r["latency"] = Value::number(lat);
  if (!std::regex_search(f[5], m, acc_re)) return fail(TOK_ERR_INVALID, "no accuracy in %s", f[5].c_str());
  r["accuracy"] = Value::number(strtod(m.str().c_str(), nullptr));
  if (lat > 1.0) return fail(TOK_ERR_INVALID, "epoch training time > 1, drop it");
  return out_json(r, out);
}

int tok_elastic_observe(tok_elastic_t* e, tok_job_t* j, double latency, int has_pending,
                        int has_failed, char** out) {
  if (!e || !j) return fail(TOK_ERR_INVALID, "elastic / job is null");
  status_of(j);  // create .status BEFORE taking pointers into the document (insertion reallocates)
  Value* specs = task_specs(j);
  Value* wspec = specs ? specs->find("Worker") : nullptr;
  if (!wspec) return fail(TOK_ERR_INVALID, "job has no Worker task to scale");
  const std::string key = job_namespace(j) + "/" + job_name(j);
  const Value* pol = j->doc.path({"spec", "torchElasticPolicy"});
  const Value* vmin = pol ? pol->find("numMinReplicas") : nullptr;
  const Value* vmax = pol ? pol->find("numMaxReplicas") : nullptr;
  // 1. policy incomplete: stop managing (elastic_scale.go:71-76)
  if (!vmin || !vmax || !vmin->is_number() || !vmax->is_number()) {
    e->metrics.erase(key);
    return result("forget", num_tasks(*wspec), nullptr, "torchjob does not configure the max or min replicas", out);
  }
  const int64_t mn = vmin->as_int(), mx = vmax->as_int();
  auto& metrics = e->metrics[key];
  const int64_t cur = num_tasks(*wspec);
  // 2. first pass: initialise the status (:79-88, job.go:41-52)
  Value* ws = worker_status(j);
  if (!ws) {
    Value s = Value::object();
    s["elasticCondition"] = Value::str("Start");
    s["continue"] = Value::boolean(true);
    s["curReplicas"] = Value::integer(cur);
    status_of(j)["elasticScalingStatues"]["Worker"] = std::move(s);
    return result("init", cur, worker_status(j), "torchelastic status initialised", out);
  }
  // 3. completed or being deleted (:92-98)
  const Value* ct = j->doc.path({"status", "completionTime"});
  const Value* dt = j->doc.path({"metadata", "deletionTimestamp"});
  if ((ct && !ct->is_null()) || (dt && !dt->is_null())) {
    e->metrics.erase(key);
    return result("forget", cur, ws, "torchjob has already completed (or been deleted)", out);
  }
  auto set = [&](const char* cond, bool cont, int64_t cur_r, int64_t last_r, const char* msg,
                 bool touch_replicas) {
    (*ws)["elasticCondition"] = Value::str(cond);
    (*ws)["continue"] = Value::boolean(cont);
    if (touch_replicas) {
      (*ws)["curReplicas"] = Value::integer(cur_r);
      (*ws)["lastReplicas"] = Value::integer(last_r);
    }
    (*ws)["message"] = Value::str(msg);
  };
  const int64_t last = ws->find("lastReplicas") ? ws->find("lastReplicas")->as_int() : 0;
  // 4. pending replicas while above the minimum: go back to the last size (:107-122, job.go:54-63)
  if (has_pending && cur > mn) {
    (*wspec)["numTasks"] = Value::integer(last);
    const int64_t prev_cur = ws->find("curReplicas") ? ws->find("curReplicas")->as_int() : cur;
    set("Stop", false, last, prev_cur, "There exists pending pods, return to the last replicas", true);
    return result("revert", last, ws, "pending replicas above the minimum", out);
  }
  // 5. cannot even hold the minimum, or a replica failed: stop managing (:123-131)
  if ((has_pending && cur == mn) || has_failed)
    return result("stop_managing", cur, ws,
                  "pods reach to running state is less than the min replicas settled, or exists pod failed", out);
  // 6./7. scaling already finished (:133-164)
  const bool cont = ws->find("continue") ? ws->find("continue")->as_bool() : false;
  const std::string cond = ws->find("elasticCondition") ? ws->find("elasticCondition")->as_string() : "";
  if (!has_pending && !cont) {
    if (cond == "ReachMaxMetric") {
      (*ws)["elasticCondition"] = Value::str("Stop");
      return result("restart_stale", cur, ws, "re-form the peer group at the reverted size", out);
    }
    if (cond == "Stop" || cond == "ReachMaxReplicas")
      return result("none", cur, ws, "stop scaling because torchelastic condition is stop or the maximum replica is reached", out);
  }
  // 8. no usable observation this tick (:168-172; latency > 1 s is dropped by the parser)
  if (latency < 0 || latency > 1.0) return result("skip", cur, ws, "no torchelastic observation this tick", out);
  // 9. collect (:178-186)
  metrics[cur].push_back(latency);
  const int n = static_cast<int>(metrics[cur].size());
  if (n < e->metric_count) return result("wait", cur, ws, "collecting observations", out);
  const size_t k = static_cast<size_t>(e->metric_count - 1);
  auto scale_up = [&]() {
    int64_t next = cur * 2;          // computeNewReplicas, job.go:102-104
    if (next > mx) next = mx;        // clamp (the reference can exceed numMaxReplicas)
    (*wspec)["numTasks"] = Value::integer(next);
    set("Continue", true, next, cur, "Pytorch job continues to be scaled", true);
    metrics[next];                   // make sure the bucket exists
    return result("scale", next, ws, "scale out", out);
  };
  if (cur > mn && cur <= mx) {
    auto lit = metrics.find(last);
    const bool have_last = lit != metrics.end() && lit->second.size() > k;
    // IsSatisfyElasticContinue (job.go:94-100); with no baseline at `last` (the reference would
    // index a nil slice) treat the step as an improvement
    const bool better = !have_last ||
                        lit->second[k] / static_cast<double>(last) > metrics[cur][k] / static_cast<double>(cur);
    if (better) {
      if (cur == mx) {  // 10. (:191-193)
        set("ReachMaxReplicas", false, 0, 0, "Pytorch job has reached the max replicas", false);
        metrics[cur].clear();
        return result("none", cur, ws, "reached numMaxReplicas", out);
      }
      return scale_up();  // 11. (:194-204)
    }
    // 12. not better: revert (:205-213, job.go:84-92)
    (*wspec)["numTasks"] = Value::integer(last);
    set("ReachMaxMetric", false, last, cur, "Pytorch job has reached the max metrics", true);
    metrics[last].clear();
    metrics[cur].clear();
    return result("revert", last, ws, "latency per replica did not improve", out);
  }
  if (cur == mn && cur < mx) return scale_up();  // 13. (:214-225)
  if (cur == mx) {                               // 14. (:227-232)
    set("ReachMaxReplicas", false, 0, 0, "Pytorch job has reached the max replicas", false);
    metrics[cur].clear();
    return result("none", cur, ws, "reached numMaxReplicas", out);
  }
  return result("none", cur, ws, "replica count outside [min, max]", out);
}

}  // extern "C"
