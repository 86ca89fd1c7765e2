/*
 * TEST INFRASTRUCTURE (oracle) — not product code. PARITY UNPINNED against real spconv (absent, un-pinned
 * third-party dependency: `spconv-cu113` 2.x + `cumm`, doc/INSTALL.md:27 of the reference).
 *
 * Sequential CPU restatement of the published SECOND / spconv `points_to_voxel` algorithm that
 * spconv.utils.Point2VoxelCPU3d.point_to_voxel implements, as called by the reference at
 *   opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:59-72 (construction: vsize_xyz, coors_range_xyz,
 *   max_num_points_per_voxel=32, num_point_features=4, max_num_voxels) and :96-116 (call + output dict
 *   voxel_features [M,32,4] zero padded, voxel_coords [M,3] int32 (z,y,x), voxel_num_points [M] int32).
 *
 * Semantics: iterate points in input order; c_j = floorf((p_j - lo_j) / vs_j) in fp32 (true division);
 * drop the point if any c_j is outside [0, grid_j); first-come voxel ids up to max_voxels; first-come
 * <= max_points points per voxel. grid_j = round((hi_j - lo_j) / vs_j).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int a2x_oracle_voxelize(const float* points, int num_points, int num_features, const float* range6,
                        const float* vsize3, int max_points, int max_voxels, float* voxels /*[max_voxels][max_points][F]*/,
                        int32_t* coords /*[max_voxels][3] zyx*/, int32_t* num_per_voxel /*[max_voxels]*/,
                        int32_t* point_voxel /*[num_points] voxel id or -1 (dropped); may be NULL*/) {
    int grid[3];
    for (int j = 0; j < 3; ++j) grid[j] = (int)roundf((range6[3 + j] - range6[j]) / vsize3[j]);
    const long long cells = (long long)grid[0] * grid[1] * grid[2];
    int32_t* table = (int32_t*)malloc(sizeof(int32_t) * (size_t)cells);
    if (!table) return -1;
    for (long long i = 0; i < cells; ++i) table[i] = -1;
    int voxel_num = 0;
    for (int i = 0; i < num_points; ++i) {
        const float* p = points + (size_t)i * num_features;
        int c[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            const float q = floorf((p[j] - range6[j]) / vsize3[j]);
            if (!(q >= 0.0f) || !(q < (float)grid[j])) {
                ok = 0;
                break;
            }
            c[j] = (int)q;
        }
        if (point_voxel) point_voxel[i] = -1;
        if (!ok) continue;
        const long long cell = ((long long)c[2] * grid[1] + c[1]) * grid[0] + c[0];
        int v = table[cell];
        if (v == -1) {
            if (voxel_num >= max_voxels) continue;
            v = voxel_num++;
            table[cell] = v;
            coords[3 * v + 0] = c[2];
            coords[3 * v + 1] = c[1];
            coords[3 * v + 2] = c[0];
            num_per_voxel[v] = 0;
        }
        if (num_per_voxel[v] < max_points) {
            memcpy(voxels + ((size_t)v * max_points + num_per_voxel[v]) * num_features, p, sizeof(float) * num_features);
            num_per_voxel[v] += 1;
            if (point_voxel) point_voxel[i] = v;
        }
    }
    free(table);
    return voxel_num;
}
