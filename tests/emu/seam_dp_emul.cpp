// tests/emu/seam_dp_emul.cpp -- TEST INFRASTRUCTURE ONLY.
// Host build of the marked DP region of imagestitch_b200/csrc/seam.cu (DpArgs, k_seam_dp<LPT>: forward pass over the TMA ring,
// reachability, back-track) on the multi-threaded block emulator.
#include "tma.cuh"

namespace is {
#include "seam_dp_region.inc"
}
using namespace is;

// P, Q: [steps][pitch] (pitch = nt * lpt); control: [steps][pitch] scratch; seam_lane: [steps + 1]
extern "C" int emu_seam_dp(const float* P, const float* Q, int lanes, int pitch, int steps, int s0, int lane0, int s1, int lane1, int lpt, int G, int D,
                           uint8_t* control, int* seam_lane, int* reached) {
    DpArgs A;
    A.P = P; A.Q = Q; A.control = control;
    A.lanes = lanes; A.pitch = pitch; A.steps = steps;
    A.s0 = s0; A.lane0 = lane0; A.s1 = s1; A.lane1 = lane1;
    A.seam_lane = seam_lane; A.reached = reached;
    A.G = G; A.D = D;
    const unsigned nt = (unsigned)(pitch / lpt);
    if ((size_t)D * 2 * G * pitch * sizeof(float) + (size_t)D * 8 > sizeof(emu_dynamic_smem)) return -1;
    if (lpt == 4) emu_launch_mt(1, nt, [&] { k_seam_dp<4>(A); });
    else if (lpt == 8) emu_launch_mt(1, nt, [&] { k_seam_dp<8>(A); });
    else if (lpt == 16) emu_launch_mt(1, nt, [&] { k_seam_dp<16>(A); });
    else return -2;
    return 0;
}
