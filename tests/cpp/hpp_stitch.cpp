// A reference-style main() over include/imagestitch.hpp: warp -> DP seam -> multi-band blend of a small synthetic strip, stage by
// stage through the C++ classes (RotationWarper / DpSeamFinder / MultiBandBlender), host buffers only.
//   hpp_stitch IN.bin OUT.bin
// IN : int32 n, h, w; float scale; per image float K[9], R[9]; per image u8 BGR[h*w*3]
// OUT: int32 roi[4]; per image int32 tl[2], size[2] + u8 seam mask; s16 pano[H*W*3]; u8 pano_mask[H*W]
#include <cstdint>
#include <cstdio>
#include <vector>

#include "imagestitch.hpp"

static is_mat host_mat(void* p, int rows, int cols, int ch, int depth, size_t elem) {
    is_mat m;
    m.data = p; m.rows = rows; m.cols = cols; m.channels = ch; m.depth = depth; m.step = (size_t)cols * ch * elem; m.device = -1;
    return m;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int32_t hdr[3];
    float scale;
    if (std::fread(hdr, 4, 3, f) != 3 || std::fread(&scale, 4, 1, f) != 1) return 4;
    const int n = hdr[0], h = hdr[1], w = hdr[2];
    std::vector<float> K(9 * (size_t)n), R(9 * (size_t)n);
    for (int i = 0; i < n; ++i)
        if (std::fread(&K[9 * (size_t)i], 4, 9, f) != 9 || std::fread(&R[9 * (size_t)i], 4, 9, f) != 9) return 4;
    std::vector<std::vector<uint8_t>> src((size_t)n, std::vector<uint8_t>((size_t)h * w * 3));
    for (int i = 0; i < n; ++i)
        if (std::fread(src[(size_t)i].data(), 1, src[(size_t)i].size(), f) != src[(size_t)i].size()) return 4;
    std::fclose(f);
    try {
        is::Context ctx(0);
        is::RotationWarper warper(ctx, IS_PROJ_CYLINDRICAL, scale);
        std::vector<is_point> corners((size_t)n);
        std::vector<is_size> sizes((size_t)n);
        std::vector<std::vector<uint8_t>> wimg((size_t)n), wmask((size_t)n);
        std::vector<is_mat> images_warped((size_t)n), masks_warped((size_t)n);
        for (int i = 0; i < n; ++i) {
            is_mat s = host_mat(src[(size_t)i].data(), h, w, 3, IS_8U, 1);
            warper.warpRoi(is_size{w, h}, &K[9 * (size_t)i], &R[9 * (size_t)i], &sizes[(size_t)i]);
            wimg[(size_t)i].resize((size_t)sizes[(size_t)i].width * sizes[(size_t)i].height * 3);
            wmask[(size_t)i].resize((size_t)sizes[(size_t)i].width * sizes[(size_t)i].height);
            images_warped[(size_t)i] = host_mat(wimg[(size_t)i].data(), sizes[(size_t)i].height, sizes[(size_t)i].width, 3, IS_8U, 1);
            masks_warped[(size_t)i] = host_mat(wmask[(size_t)i].data(), sizes[(size_t)i].height, sizes[(size_t)i].width, 1, IS_8U, 1);
            corners[(size_t)i] = warper.warpWithMask(s, &K[9 * (size_t)i], &R[9 * (size_t)i], images_warped[(size_t)i], masks_warped[(size_t)i]);
        }
        is::DpSeamFinder(ctx, IS_COST_COLOR).find(images_warped, corners, masks_warped);
        is::MultiBandBlender blender(ctx, 0, 3, IS_WEIGHT_32F);
        blender.prepare(corners, sizes);
        for (int i = 0; i < n; ++i) blender.feed(images_warped[(size_t)i], masks_warped[(size_t)i], corners[(size_t)i]);
        const is_size ds = blender.dstSize();
        std::vector<int16_t> pano((size_t)ds.width * ds.height * 3);
        std::vector<uint8_t> pmask((size_t)ds.width * ds.height);
        is_mat pm = host_mat(pano.data(), ds.height, ds.width, 3, IS_16S, 2), mm = host_mat(pmask.data(), ds.height, ds.width, 1, IS_8U, 1);
        blender.blend(pm, mm);
        ctx.synchronize();
        FILE* o = std::fopen(argv[2], "wb");
        if (!o) return 5;
        int minx = corners[0].x, miny = corners[0].y;
        for (int i = 1; i < n; ++i) { if (corners[(size_t)i].x < minx) minx = corners[(size_t)i].x; if (corners[(size_t)i].y < miny) miny = corners[(size_t)i].y; }
        const int32_t roi[4] = {minx, miny, ds.width, ds.height};
        std::fwrite(roi, 4, 4, o);
        for (int i = 0; i < n; ++i) {
            const int32_t g[4] = {corners[(size_t)i].x, corners[(size_t)i].y, sizes[(size_t)i].width, sizes[(size_t)i].height};
            std::fwrite(g, 4, 4, o);
            std::fwrite(wmask[(size_t)i].data(), 1, wmask[(size_t)i].size(), o);
        }
        std::fwrite(pano.data(), 2, pano.size(), o);
        std::fwrite(pmask.data(), 1, pmask.size(), o);
        std::fclose(o);
    } catch (const is::Error& e) {
        std::fprintf(stderr, "is::Error %d: %s\n", e.status, e.what());
        return 10;
    }
    return 0;
}
