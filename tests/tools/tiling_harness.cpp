// Test harness (CPU): exposes monocularsfm_b200/csrc/ba_tiles.hpp through a tiny C interface so that
// tests/test_ba_tiles.py can check the structure analysis without a GPU.
#include <cstdlib>
#include <cstring>
#include "../../monocularsfm_b200/csrc/ba_tiles.hpp"

using namespace msfm::ba;
static Tiling g_T;

extern "C" {
// returns 0 on success, 1 if a point is observed twice by one camera; sizes[] = n_tiles, n_tile_cams, n_marks, n_blocks, w_max, n_items, first_long, n_long
int tiling_build(int n_cams, int n_pts, int n_obs, const int32_t* obs_cam, const int32_t* obs_pt, const int32_t* cam_free,
                 int n_free, int max_obs, int max_pts, int max_items, int32_t sizes[9]) {
    TilingParams p; p.max_obs = max_obs; p.max_pts = max_pts; p.max_items = max_items;
    if (const char* e = std::getenv("TILING_TEST_CHUNK_PTS")) p.chunk_pts = std::atoi(e);      // tests: tiny chunks exercise the concatenation
    if (const char* e = std::getenv("TILING_TEST_CHUNK_LONG")) p.chunk_long = std::atoi(e);
    g_T = Tiling();
    if (!build_tiling(n_cams, n_pts, n_obs, obs_cam, obs_pt, cam_free, p, g_T)) return 1;
    std::vector<uint8_t> present;
    mark_blocks(g_T, cam_free, n_free, present);
    assign_slots(g_T, cam_free, n_free, present);
    sizes[0] = (int32_t)g_T.tiles.size(); sizes[1] = (int32_t)g_T.tile_cams.size(); sizes[2] = (int32_t)g_T.tile_marks.size();
    sizes[3] = (int32_t)g_T.blk_col.size(); sizes[4] = g_T.w_max; sizes[5] = (int32_t)g_T.items.size(); sizes[6] = g_T.first_long; sizes[7] = g_T.n_long; sizes[8] = (int32_t)g_T.runs.size();
    return 0;
}
void tiling_fetch(int32_t* pt_order, int32_t* pt_start, int32_t* obs_perm, uint8_t* obs_lcam, uint8_t* obs_lpt, int32_t* tiles /*[n][12]*/,
                  int32_t* items /*[n][12]*/, int32_t* tile_cams, int32_t* tile_slots, int32_t* blk_row, int32_t* blk_col, uint32_t* runs) {
    std::memcpy(pt_order, g_T.pt_order.data(), g_T.pt_order.size() * 4);
    std::memcpy(pt_start, g_T.pt_start.data(), g_T.pt_start.size() * 4);
    std::memcpy(obs_perm, g_T.obs_perm.data(), g_T.obs_perm.size() * 4);
    std::memcpy(obs_lcam, g_T.obs_lcam.data(), g_T.obs_lcam.size());
    std::memcpy(obs_lpt, g_T.obs_lpt.data(), g_T.obs_lpt.size());
    static_assert(sizeof(Tile) == 48 && sizeof(Item) == 48, "Tile / Item layout");
    std::memcpy(tiles, g_T.tiles.data(), g_T.tiles.size() * sizeof(Tile));
    std::memcpy(items, g_T.items.data(), g_T.items.size() * sizeof(Item));
    if (runs) std::memcpy(runs, g_T.runs.data(), g_T.runs.size() * 4);
    std::memcpy(tile_cams, g_T.tile_cams.data(), g_T.tile_cams.size() * 4);
    std::memcpy(tile_slots, g_T.tile_slots.data(), g_T.tile_slots.size() * 4);
    std::memcpy(blk_row, g_T.blk_row.data(), g_T.blk_row.size() * 4);
    std::memcpy(blk_col, g_T.blk_col.data(), g_T.blk_col.size() * 4);
}
}
