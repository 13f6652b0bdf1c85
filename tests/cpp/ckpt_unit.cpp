// CPU-only check of the checkpoint file helpers behind RT_B200_CHECKPOINT (host/rtb200_host.hpp): save -> load round
// trip, atomic replacement, and the refusal of a file written for another render.  usage: ckpt_unit <file> <mode>
#include "common/rtweekend.hpp"
#include "core/camera.hpp"

int main(int argc, char* argv[]) {
  if (argc < 3) return 2;
  const char* path = argv[1];
  const std::string mode = argv[2];
  rtb200::checkpoint_header h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "RTB2CKPT", 8);
  h.version = 1, h.width = 5, h.height = 3, h.spp_total = 100, h.max_depth = 7, h.seed = 0, h.scene_hash = 0x1234567890abcdefULL;
  std::vector<int64_t> sums(5 * 3 * 3), back(5 * 3 * 3, -1);
  for (size_t i = 0; i < sums.size(); i++) sums[i] = int64_t(i) * 1000003 - 7;
  if (mode == "roundtrip") {
    if (rtb200::checkpoint_load(path, h, back) != 0) return 10;  // no file yet
    rtb200::checkpoint_save(path, h, 40, sums);
    if (rtb200::checkpoint_load(path, h, back) != 40 || back != sums) return 11;
    sums[3] = 99;
    rtb200::checkpoint_save(path, h, 60, sums);  // replaces the file
    if (rtb200::checkpoint_load(path, h, back) != 60 || back[3] != 99) return 12;
    return 0;
  }
  if (mode == "other_scene") h.scene_hash ^= 1;  // exits 1 with a message
  if (mode == "other_spp") h.spp_total = 101;
  rtb200::checkpoint_load(path, h, back);
  return 0;
}
