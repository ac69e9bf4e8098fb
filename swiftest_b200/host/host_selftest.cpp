// host_selftest.cpp -- drives the C++ host mirror (swiftest_host.hpp) the way symba_step_system does for one step:
//   pl%encounter_check, tp%encounter_check, pl%accel (all pairs minus the encounter pairs), tp%accel_int, pl%drift.
// Input/output are raw little-endian binaries written/read by tests/test_gpu_host_cpp.py, which compares the output
// with the CPU oracle.  usage: host_selftest in.bin out.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "swiftest_host.hpp"

using namespace swiftest;

template <class T> static void rd(FILE *f, T *p, size_t n)
{
    if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <class T> static void wr(FILE *f, const T *p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 2;
    I4B hdr[4];  // npl, ntp, flat, lmtiny
    DP par[3];   // dt, GMTINY, cb Gmass
    rd(fi, hdr, 4);
    rd(fi, par, 3);
    const I4B npl = hdr[0], ntp = hdr[1];
    swiftest_parameters param;
    param.lflatten_interactions = hdr[2] != 0;
    param.lmtiny_pl = hdr[3] != 0;
    param.GMTINY = par[1];
    param.lclose = true;
    swiftest_cb cb;
    cb.Gmass = par[2];
    const DP dt = par[0];

    symba_pl pl;
    symba_tp tp;
    pl.setup(npl);
    tp.setup(ntp);
    rd(fi, pl.rh.data(), 3 * (size_t)npl);
    rd(fi, pl.vb.data(), 3 * (size_t)npl);
    rd(fi, pl.Gmass.data(), (size_t)npl);
    rd(fi, pl.radius.data(), (size_t)npl);
    rd(fi, pl.rhill.data(), (size_t)npl);
    rd(fi, tp.rh.data(), 3 * (size_t)ntp);
    rd(fi, tp.vb.data(), 3 * (size_t)ntp);
    fclose(fi);
    pl.flatten(param);

    try {
        cuda_context c(0);
        encounter_list plpl, pltp;
        const bool lpl = pl.encounter_check(c, param, dt, 0, plpl);
        const bool ltp = tp.encounter_check(c, param, pl, dt, 0, pltp);
        pl.kick_getacch(c, param, plpl);                                      // ah starts at zero (helio_kick.f90:113)
        tp.accel_int(c, param, pl.Gmass.data(), pl.rh.data(), pl.nbody);
        const int lost = pl.helio_drift(c, cb, param, dt);

        FILE *fo = fopen(argv[2], "wb");
        I8B counts[4] = {plpl.nenc, pltp.nenc, (I8B)pl.nplm, (I8B)lost + 10 * (I8B)lpl + 100 * (I8B)ltp};
        wr(fo, counts, 4);
        wr(fo, plpl.index1.data(), (size_t)plpl.nenc);
        wr(fo, plpl.index2.data(), (size_t)plpl.nenc);
        wr(fo, pltp.index1.data(), (size_t)pltp.nenc);
        wr(fo, pltp.index2.data(), (size_t)pltp.nenc);
        wr(fo, pl.ah.data(), 3 * (size_t)npl);
        wr(fo, tp.ah.data(), 3 * (size_t)ntp);
        wr(fo, pl.rh.data(), 3 * (size_t)npl);
        wr(fo, pl.vb.data(), 3 * (size_t)npl);
        fclose(fo);
    } catch (const fatal_error &e) {
        fprintf(stderr, "%s\n", e.what());  // base_util_exit(FAILURE)
        return 1;
    }
    return 0;
}
