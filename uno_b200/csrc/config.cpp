#include "config.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace uno {
namespace {

struct Entry { const char* name; const char* env; int def; bool env_inverts; };
// name, environment variable that seeds it, default, whether the variable means "off"
const Entry kEntries[CFG_COUNT] = {
    {"tc", "UNO_B200_DISABLE_TC", 1, true},
    {"mid_tc", "UNO_B200_MID_TC", 1, false},
    {"cmm_tc", "UNO_B200_CMM_TC", 1, false},
    {"kpipe_align", "UNO_B200_KPIPE_ALIGN", 1, false},
    {"kpipe_lw16", "UNO_B200_KPIPE_LW16", 1, false},
    {"rowgemm_epi16", "UNO_B200_ROWGEMM_EPI16", 1, false},
    {"rowgemm_parity", "UNO_B200_ROWGEMM_PARITY", 1, false},
    {"norm_big_cluster", "UNO_B200_NORM_BIG_CLUSTER", 1, false},
    {"overlap", "UNO_B200_OVERLAP", 1, false},
    {"pointwise3d_fixed", "UNO_B200_POINTWISE3D_FIXED", 0, false},
    {"proj_simt", "UNO_B200_PROJ_SIMT", 0, false},
    {"nvtx", "UNO_B200_NVTX", 0, false},
    {"exp0", "UNO_B200_EXP0", 0, false},
    {"exp1", "UNO_B200_EXP1", 0, false},
    {"exp2", "UNO_B200_EXP2", 0, false},
    {"kpipe_debug", "UNO_B200_KPIPE_DEBUG", 0, false},
    {"wgrad_debug", "UNO_B200_WGRAD_DEBUG", 0, false},
};

std::atomic<int> g_val[CFG_COUNT];
std::once_flag g_once;

void init() {
    for (int k = 0; k < CFG_COUNT; ++k) {
        int v = kEntries[k].def;
        const char* e = getenv(kEntries[k].env);
        if (e && e[0]) {
            const int n = atoi(e);
            v = kEntries[k].env_inverts ? (n ? 0 : 1) : n;
        }
        g_val[k].store(v);
    }
}

int find(const char* name) {
    if (!name) return -1;
    for (int k = 0; k < CFG_COUNT; ++k)
        if (strcmp(name, kEntries[k].name) == 0) return k;
    return -1;
}

}  // namespace

int cfg(CfgKey k) {
    std::call_once(g_once, init);
    return g_val[k].load(std::memory_order_relaxed);
}

int cfg_set(const char* name, int value) {
    std::call_once(g_once, init);
    const int k = find(name);
    if (k < 0) return -1;
    g_val[k].store(value);
    return 0;
}

int cfg_get(const char* name, int* value) {
    std::call_once(g_once, init);
    const int k = find(name);
    if (k < 0) return -1;
    if (value) *value = g_val[k].load();
    return 0;
}

const char* cfg_name(int k) { return (k >= 0 && k < CFG_COUNT) ? kEntries[k].name : nullptr; }

}  // namespace uno
