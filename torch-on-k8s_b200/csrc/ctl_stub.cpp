// TEMPORARY: placeholders so that libtok8s.so exports the full ABI while the C++ control plane is
// being written; replaced by ctl_*.cpp.
#include "tok_internal.h"
#define STUB(name, ...) int name(__VA_ARGS__) { return tok::fail(TOK_ERR_UNSUPPORTED, #name ": not built yet"); }
extern "C" {
static unsigned g_gates = TOK_GATES_DEFAULT;
int tok_set_feature_gates(unsigned g) { g_gates = g; return TOK_OK; }
unsigned tok_get_feature_gates(void) { return g_gates; }
STUB(tok_job_parse, const char*, tok_job_t**)
STUB(tok_job_default, tok_job_t*)
STUB(tok_job_to_json, const tok_job_t*, char**)
void tok_job_free(tok_job_t*) {}
STUB(tok_job_cluster_spec, const tok_job_t*, const char*, int, char**)
STUB(tok_job_dag_ready, const tok_job_t*, const char*, const char*, int*)
STUB(tok_gang_admit, const tok_job_t*, int, char**)
STUB(tok_failover_decide, const char*, int, const char*, int*)
STUB(tok_job_update_status, tok_job_t*, const char*, int, const char*, char**)
STUB(tok_coord_create, int, int, uint64_t, tok_coord_t**)
void tok_coord_destroy(tok_coord_t*) {}
STUB(tok_coord_set_quota, tok_coord_t*, const char*, int)
STUB(tok_coord_set_used, tok_coord_t*, const char*, int)
STUB(tok_coord_enqueue, tok_coord_t*, const tok_job_t*, const char*)
STUB(tok_coord_is_queuing, tok_coord_t*, const char*, int*)
STUB(tok_coord_dequeue, tok_coord_t*, const char*)
STUB(tok_coord_job_settled, tok_coord_t*, const char*)
STUB(tok_coord_tick, tok_coord_t*, double, char**)
STUB(tok_coord_pending, tok_coord_t*, const char*, int*)
STUB(tok_elastic_create, int, tok_elastic_t**)
void tok_elastic_destroy(tok_elastic_t*) {}
STUB(tok_elastic_parse_log, const char*, char**)
STUB(tok_elastic_observe, tok_elastic_t*, tok_job_t*, double, int, int, char**)
}
