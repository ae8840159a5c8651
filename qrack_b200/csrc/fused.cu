// fused.cu — placeholder until the fused multi-gate sweep lands: every gate runs unfused.
#include "sv_common.cuh"
namespace b200sv {
bool fused_accepts(const State*, const GateOp&) { return false; }
int fused_flush(State* s) { s->queue.clear(); return B200SV_OK; }
}
