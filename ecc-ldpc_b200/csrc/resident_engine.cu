// resident_engine.cu -- placeholder until the on-chip engine lands (next commit).
#include "common.cuh"
namespace ldpc {
int resident_kind_for(const ldpc_b200_code *) { return 0; }
bool resident_supports(const ldpc_b200_decoder *) { return false; }
int resident_decode(ldpc_b200_decoder *, int, int, int, const void *, bool, uint32_t *, int32_t *, uint8_t *)
{ set_error("RESIDENT engine not built"); return LDPC_B200_EUNSUPPORTED; }
int resident_simulate(ldpc_b200_decoder *, int, int, float, float, int, int, uint64_t, uint64_t, uint64_t, uint64_t *)
{ set_error("RESIDENT engine not built"); return LDPC_B200_EUNSUPPORTED; }
}
