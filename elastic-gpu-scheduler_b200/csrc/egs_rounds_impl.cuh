// egs_rounds_impl.cuh -- EGS_MODE_ROUNDS (placeholder: routed to the per-pod pass until the
// round kernels land).
#pragma once

static int batch_rounds(egs_handle *h, int P, const int32_t *c_off, const egs_unit *units,
                        const std::vector<int> &slots, PodOut out) {
  return batch_rescan(h, P, c_off, units, slots, out);
}
static int rounds_sync_rows(egs_handle *) { return EGS_OK; }
static void rounds_free(RoundsState *) {}
static int rounds_comm_unique_id(uint8_t *) { return EGS_ERR_COMM; }
static int rounds_comm_init(egs_handle *, const uint8_t *) { return EGS_ERR_COMM; }
