// All-pairs Chamfer matrix (attacker/prepare_indices_for_attack.py:104-139).
#include "nn_tiles.cuh"

extern "C" int ga_chamfer_all_pairs(int s, int n, const float* clouds, int row0, int rows, float* out, int mode,
                                    ga_stream_t stream) {
  (void)s; (void)n; (void)clouds; (void)row0; (void)rows; (void)out; (void)mode; (void)stream;
  ga::set_error("ga_chamfer_all_pairs: not built yet");
  return GA_ERR_UNSUPPORTED;
}
