// egs_rounds.cuh -- EGS_MODE_ROUNDS: declarations (definitions in egs_rounds_impl.cuh,
// included at the end of egs_api.cu once egs_handle is complete).
#pragma once
#include <vector>
#include "egs_kernels.cuh"

struct egs_handle;

struct RoundsState {
  bool index_valid = false;     // per-shape candidate index is in step with rows + option tables
  void *comm = nullptr;         // ncclComm_t
};

static int batch_rescan(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out);
static int batch_rounds(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out);
static int rounds_sync_rows(egs_handle *h);
static void rounds_free(RoundsState *r);
static int rounds_comm_unique_id(uint8_t out_id[128]);
static int rounds_comm_init(egs_handle *h, const uint8_t id[128]);
