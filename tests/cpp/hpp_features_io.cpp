// The front of a reference-style main() over include/imagestitch.hpp: imread -> OrbFeaturesFinder -> (remap as a stand-alone
// call) -> imwrite, host buffers only.
//   hpp_features_io IN.bmp OUT.bmp OUT.bin
// OUT.bmp: the image written back as it was read; OUT.bin: int32 n; per key point 6 floats (x, y, size, angle, response, octave);
// n x 32 descriptor bytes.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "imagestitch.hpp"

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    try {
        is::Context ctx(0);
        const is_size sz = is::bmpSize(ctx, argv[1]);
        std::vector<uint8_t> img((size_t)sz.width * sz.height * 3);
        is_mat m;
        m.data = img.data(); m.rows = sz.height; m.cols = sz.width; m.channels = 3; m.depth = IS_8U; m.step = (size_t)sz.width * 3; m.device = -1;
        is::imread(ctx, argv[1], m);
        is::OrbFeaturesFinder finder(ctx);                   // grid 3 x 1, 1500 -> 510 features per cell, 1.3f, 5 levels ([FEAT]:39-55)
        is::ImageFeatures features;
        finder(m, features);
        std::vector<float> xmap((size_t)sz.width * sz.height), ymap(xmap.size());
        for (int y = 0; y < sz.height; ++y)
            for (int x = 0; x < sz.width; ++x) { xmap[(size_t)y * sz.width + x] = (float)x; ymap[(size_t)y * sz.width + x] = (float)y; }
        is_mat mx = m, my = m, out = m;
        std::vector<uint8_t> same(img.size());
        mx.data = xmap.data(); mx.channels = 1; mx.depth = IS_32F; mx.step = (size_t)sz.width * 4;
        my = mx; my.data = ymap.data();
        out.data = same.data();
        is::RotationWarper::remap(ctx, m, mx, my, IS_INTER_LINEAR, IS_BORDER_REFLECT, out);      // identity maps: `same` == img
        is::imwrite(ctx, argv[2], out);
        FILE* f = std::fopen(argv[3], "wb");
        if (!f) return 3;
        const int32_t n = (int32_t)features.keypoints.size();
        std::fwrite(&n, 4, 1, f);
        for (const is_keypoint& k : features.keypoints) {
            const float v[6] = {k.x, k.y, k.size, k.angle, k.response, (float)k.octave};
            std::fwrite(v, 4, 6, f);
        }
        std::fwrite(features.descriptors.data(), 1, features.descriptors.size(), f);
        std::fclose(f);
    } catch (const is::Error& e) {
        std::fprintf(stderr, "is::Error %d: %s\n", e.status, e.what());
        return 10;
    }
    return 0;
}
