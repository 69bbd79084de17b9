// NCCL tuner plugin (ncclTunerPlugin_v3 / _v4), packaged inside the net plugin library.
//
// Why: over this transport the right protocol depends on the message size in a way NCCL's built-in model cannot know.
//   * LL keeps its buffers in host memory (they ride our shared-memory ring): best latency for tiny messages, slow for
//     anything larger (measured at 2 ranks: 26-44 us up to 8 KiB, but 124-395 us at 32-512 KiB).
//   * Simple rides the NVLink kernels: ~63 us floor, then bandwidth.
//   * LL128 relies on 128-byte store atomicity between the data and its flag, which a copy kernel moving 16-byte vectors
//     in arbitrary order does not preserve: never offered.
// So: LL up to BNET_TUNER_LL_MAX bytes (default 8192), Simple above, LL128 never.  The tuner only speaks when the bnet
// engine is the active network with its device path on (BNET_TUNER=0 silences it); otherwise the cost table is left alone.
//
// The reference has no tuner (NCCL had no such plugin when it was written); its README tunes by hand through environment
// variables (reference README.md:20-46).
#include <stdlib.h>
#include <string.h>

#include "bnet/nccl_net_abi.h"
#include "core/engine.h"

using namespace bnet;

#define BNET_EXPORT __attribute__((visibility("default")))

namespace {

constexpr float kIgnore = -1.0f;            // NCCL_ALGO_PROTO_IGNORE
enum { PROTO_LL = 0, PROTO_LL128 = 1, PROTO_SIMPLE = 2 };

struct TunerCtx {
  bool active;
  size_t ll_max;
  bool prefer_collnet;     // BNET_COLLNET=1 (the CollNet table is on): make NCCL take its CollNet algorithms wherever it offers them
};
enum { ALGO_COLLNET_DIRECT = 2, ALGO_COLLNET_CHAIN = 3 };   // NCCL_ALGO_* (tree 0, ring 1, collnet direct 2, collnet chain 3, nvls 4, ...)

ncclResult_t tuner_init(size_t nRanks, size_t nNodes, ncclDebugLogger_t logFunction, void** context) {
  (void)nRanks; (void)nNodes;
  if (logFunction) log_set_nccl_logger(logFunction);
  TunerCtx* c = new TunerCtx();
  const Config& cfg = Config::get();
  // Only when NCCL's traffic actually rides our device path.  (NCCL may have loaded this library a second time under the
  // tuner's file name — a separate instance with its own engine singleton — so the verdict is derived from scratch here:
  // same configuration, same CUDA probe, and bnet selected as the network.)
  const char* net = getenv("NCCL_NET");
  const char* netp = getenv("NCCL_NET_PLUGIN");
  const bool ours = (net && strstr(net, "BNet")) || (netp && strstr(netp, "bnet")) || Engine::get().cuda_ok();
  Engine::get().init();
  c->active = env_int("TUNER", 1) != 0 && cfg.nvl && cfg.gdr && ours && Engine::get().cuda_ok();
  c->ll_max = (size_t)env_int("TUNER_LL_MAX", 8192);
  c->prefer_collnet = env_int("TUNER_PREFER_COLLNET", env_int("COLLNET", 0)) != 0;
  if (c->active) {
    // the protocol choice is ours now: lift the blanket "Simple only" default the net plugin's init supplied (if it was us)
    const char* p = getenv("NCCL_PROTO");
    if (p && !strcmp(p, "Simple") && getenv("BNET_PROTO_DEFAULTED")) setenv("NCCL_PROTO", "LL,Simple", 1);
    BNET_INFO("tuner: LL up to %zu bytes, Simple above, LL128 never (BNET_TUNER=0 disables)", c->ll_max);
  }
  *context = c;
  return ncclSuccess;
}

ncclResult_t tuner_coll_info(void* context, size_t nBytes, float** collCostTable, int numAlgo, int numProto) {
  TunerCtx* c = static_cast<TunerCtx*>(context);
  if (!c || !c->active || !collCostTable || numProto <= PROTO_SIMPLE) return ncclSuccess;
  float* table = reinterpret_cast<float*>(collCostTable);     // float[numAlgo][numProto]
  for (int a = 0; a < numAlgo; a++) {
    float* row = table + (size_t)a * numProto;
    row[PROTO_LL128] = kIgnore;
    if (nBytes <= c->ll_max) {
      if (row[PROTO_LL] != kIgnore) row[PROTO_SIMPLE] = kIgnore;     // LL where NCCL offers it
    } else {
      if (row[PROTO_SIMPLE] != kIgnore) row[PROTO_LL] = kIgnore;
    }
  }
  // With the CollNet table switched on the user wants the offload: where NCCL offers a CollNet algorithm for this
  // collective (entry not "ignore": the table supports the type / op and the topology allows it), it wins.
  if (c->prefer_collnet && numAlgo > ALGO_COLLNET_CHAIN)
    for (int a : {ALGO_COLLNET_DIRECT, ALGO_COLLNET_CHAIN}) {
      float* row = table + (size_t)a * numProto;
      if (row[PROTO_SIMPLE] != kIgnore) row[PROTO_SIMPLE] = 0.0f;
    }
  return ncclSuccess;
}

ncclResult_t tuner_v3_coll_info(void* context, int collType, size_t nBytes, int numPipeOps, float** collCostTable, int numAlgo,
                                int numProto, int* nChannels) {
  (void)collType; (void)numPipeOps; (void)nChannels;
  return tuner_coll_info(context, nBytes, collCostTable, numAlgo, numProto);
}
ncclResult_t tuner_v4_coll_info(void* context, int collType, size_t nBytes, int numPipeOps, float** collCostTable, int numAlgo,
                                int numProto, int regBuff, int* nChannels) {
  (void)collType; (void)numPipeOps; (void)regBuff; (void)nChannels;
  return tuner_coll_info(context, nBytes, collCostTable, numAlgo, numProto);
}

ncclResult_t tuner_destroy(void* context) {
  delete static_cast<TunerCtx*>(context);
  return ncclSuccess;
}

}  // namespace

extern "C" {
typedef struct {
  const char* name;
  ncclResult_t (*init)(size_t nRanks, size_t nNodes, ncclDebugLogger_t logFunction, void** context);
  ncclResult_t (*getCollInfo)(void* context, int collType, size_t nBytes, int numPipeOps, float** collCostTable, int numAlgo,
                              int numProto, int* nChannels);
  ncclResult_t (*destroy)(void* context);
} bnetTuner_v3_t;
typedef struct {
  const char* name;
  ncclResult_t (*init)(size_t nRanks, size_t nNodes, ncclDebugLogger_t logFunction, void** context);
  ncclResult_t (*getCollInfo)(void* context, int collType, size_t nBytes, int numPipeOps, float** collCostTable, int numAlgo,
                              int numProto, int regBuff, int* nChannels);
  ncclResult_t (*destroy)(void* context);
} bnetTuner_v4_t;

BNET_EXPORT bnetTuner_v3_t ncclTunerPlugin_v3 = {"BNet", tuner_init, tuner_v3_coll_info, tuner_destroy};
BNET_EXPORT bnetTuner_v4_t ncclTunerPlugin_v4 = {"BNet", tuner_init, tuner_v4_coll_info, tuner_destroy};
}
