// All-reduce over a full mesh of plugin connections (csrc/coll/transport_mesh.cc): the C++ face the CollNet table
// (csrc/plugin/collnet.cc) drives, next to the C API of the Python wrappers.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "core/engine.h"

struct BnetTMesh;

namespace bnet {

struct MeshMr;   // one buffer registered with every connection of a mesh
struct MeshOp;   // one all-reduce in flight: resumable, driven by tmesh_op_step()

enum MeshAlgo : int {
  MESH_ONE_SHOT = 0,   // every rank adds its whole input into every peer's output: 1 network step, (n-1) x size sent
  MESH_TWO_SHOT = 1,   // reduce-scatter into the slice owners (fused isend), then all-gather by copy: 2 steps,
                       // 2 (n-1)/n x size sent — the bandwidth-optimal shape on a switch; every rank ends with the same bits
};

// `listen` is adopted (deleted with the mesh).  Its handle has been given to every other rank out of band.
BnetTMesh* tmesh_new(ListenComm* listen, int rank, int world, int net_dev);
// `handles`: world x NCCL_NET_HANDLE_MAXSIZE bytes, entry r produced by rank r.  Blocks until the mesh is up (or timeout).
int tmesh_connect(BnetTMesh* m, const void* handles, int timeout_ms);
MeshMr* tmesh_reg(BnetTMesh* m, void* data, size_t bytes, int ptr_type);
void tmesh_dereg(BnetTMesh* m, MeshMr* mr);
bool tmesh_mr_covers(const MeshMr* mr, const void* p, size_t bytes);
// nullptr: refused (bnet_tmesh_last_error).  dtype: 0 = fp32, 1 = bf16 (pairs: equal, or bf16 in / fp32 out).
MeshOp* tmesh_op_start(BnetTMesh* m, int algo, const void* in, MeshMr* in_mr, void* out, MeshMr* out_mr, size_t count,
                       int in_dtype, int out_dtype, size_t piece_bytes, int inflight, int timeout_ms);
int tmesh_op_step(MeshOp* op);   // 1 = finished, 0 = in progress, -1 = failed (bnet_tmesh_last_error)
void tmesh_op_free(MeshOp* op);
void tmesh_destroy(BnetTMesh* m);
const char* tmesh_error(BnetTMesh* m);

}  // namespace bnet
