// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Link-time stand-ins for the handful of runtime
// symbols that the reference's bvh.cpp / mesh.cpp expect from the rest of warp.so
// (allocator, error string, APIC recorder, cuBQL host backend).  None of them is on the
// LBVH / mesh-query path; they only have to exist so the reference sources link unmodified.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "warp.h"
#include "bvh.h"

static char g_ref_error[4096];

namespace wp {
void set_error_string(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ref_error, sizeof(g_ref_error), fmt, ap);
    va_end(ap);
}
void cubql_bvh_create_host(vec3*, vec3*, int, int, BVH&) { }
void cubql_bvh_destroy_host(BVH&) { }
void cubql_bvh_refit_host(BVH&) { }
void cubql_bvh_rebuild_host(BVH&) { }
}  // namespace wp

extern "C" const char* ref_get_error_string() { return g_ref_error; }

void* wp_alloc_host(size_t s, const char*) { return malloc(s); }
void wp_free_host(void* p) { free(p); }

struct APICState;
extern "C" APICState* wp_apic_get_recording_state() { return nullptr; }
void apic_record_bvh_refit(APICState*, uint64_t) { }
void apic_record_bvh_rebuild(APICState*, uint64_t, int32_t) { }
