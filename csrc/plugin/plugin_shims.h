// Pieces of the ncclNet shims (csrc/plugin/plugin.cc) that the CollNet tables (csrc/plugin/collnet.cc) share: one init,
// one device list, the same per-version property structs.
#pragma once

#include "bnet/nccl_net_abi.h"

namespace bnet {
namespace plugin {

ncclResult_t init(ncclDebugLogger_t logfn);
ncclResult_t v4_props(int dev, ncclNetProperties_v4_t* o);
ncclResult_t v6_props(int dev, ncclNetProperties_v6_t* o);
ncclResult_t v7_props(int dev, ncclNetProperties_v7_t* o);
ncclResult_t v8_props(int dev, ncclNetProperties_v8_t* o);
ncclResult_t v9_props(int dev, ncclNetProperties_v9_t* o);

}  // namespace plugin
}  // namespace bnet
