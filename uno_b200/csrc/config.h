// Run-time switches of the library, read from the environment ONCE (first use) and changeable afterwards only through
// the C ABI (uno_config_set) -- no getenv() on any call path.  Every switch selects between kernels that have all been
// run and parity-checked on a B200 (tests/test_gpu_variants.py); the defaults are the measured-fastest configuration.
#pragma once

namespace uno {

enum CfgKey {
    CFG_TC = 0,             // tensor-core (tcgen05) kernels; 0 = fp32 SIMT kernels everywhere        env UNO_B200_DISABLE_TC=1 -> 0
    CFG_MID_TC,             // leading-axis transform on tcgen05 (tc_mid.cuh)                          env UNO_B200_MID_TC
    CFG_CMM_TC,             // per-mode channel contraction on tcgen05 (tc_cmm.cuh) where the shape fills a tile; 2 = always
    CFG_KPIPE_ALIGN,        // analysis kernel: 16-byte row-class loads for rows that are not 16-byte aligned
    CFG_KPIPE_LW16,         // analysis kernel: 16 loader warps
    CFG_ROWGEMM_EPI16,      // synthesis kernel: 16 epilogue warps
    CFG_ROWGEMM_PARITY,     // synthesis kernel: row-class tiles for row pitches that are not multiples of 8 floats (2: two-class parity tiles only, 0: none)
    CFG_NORM_BIG_CLUSTER,   // InstanceNorm cluster kernels with 200 KB per CTA for planes beyond 8 x 72 KB
    CFG_OVERLAP,            // fork / join of the two block branches (pointwise / spectral) on a library-owned side stream; default on
    CFG_POINTWISE3D_FIXED,  // NOT the reference's behaviour: band-limited 3-D pointwise resample
    CFG_PROJ_SIMT,          // projection backward on the fp32 kernel instead of the tcgen05 one
    CFG_NVTX,               // NVTX ranges per C-ABI call and per fused spectral convolution (U-level), for nsys / ncu --nvtx
    CFG_EXP0, CFG_EXP1, CFG_EXP2,   // scratch switches for kernel experiments (0 = shipped behaviour)
    CFG_KPIPE_DEBUG,        // timing probes (tools/kpipe_probe.py): results become garbage
    CFG_WGRAD_DEBUG,        // timing probes (tools/wgrad_probe.py)
    CFG_COUNT
};

int cfg(CfgKey k);                          // current value
int cfg_set(const char* name, int value);   // 0 ok, -1 unknown name
int cfg_get(const char* name, int* value);  // 0 ok, -1 unknown name
const char* cfg_name(int k);                // "tc", "mid_tc", ... (NULL past the end)

}  // namespace uno
