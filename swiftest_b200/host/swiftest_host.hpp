// swiftest_host.hpp -- C++ host-side mirror of the reference's interfaces for the force-and-drift hot path, written over
// the C ABI (include/swiftest_cuda.h).  The reference host is Modern Fortran and no Fortran compiler exists in this
// image, so this header is the compiled stand-in for the Fortran submodule bodies of fortran/swiftest_kick_cuda.f90:
// same procedure names, argument order and meaning, early returns and error behaviour as the reference
// ("file:line" relative to /root/reference/src).  Header only; link with -lswiftest_cuda.
//
//   generic swiftest_kick_getacch_int_all      swiftest/swiftest_module.f90:940-991 (5 specifics, resolved by argument type)
//   swiftest_drift_all                         swiftest/swiftest_drift.f90:60-108
//   encounter_check_all_plpl/_plplm/_pltp      encounter/encounter_check.f90:14-140
//   swiftest_pl / swiftest_tp / symba_pl       the type-bound procedures accel_int, drift, encounter_check, set_renc,
//                                              flatten (swiftest_module.f90:123-331, symba/symba_module.f90:30-70)
//
// Arrays use the Fortran memory layout: r(3,n) == std::vector<double> of size 3n {x1,y1,z1,x2,...}; indices are 1-based.
#pragma once
#include <cstdint>
#include <algorithm>
#include <stdexcept>
#include <utility>
#include <string>
#include <vector>

#include "../../include/swiftest_cuda.h"

namespace swiftest {

using I4B = int32_t;
using I8B = int64_t;
using DP = double;

// base_parameters, the switches the hot path reads (base/base_module.f90:24-135)
struct swiftest_parameters {
    bool lflatten_interactions = false;  // INTERACTION_LOOPS FLAT     (swiftest_io.f90:2707-2718)
    bool lclose = true;                  // CHK_CLOSE: radius-checked variants (kick.f90:29,35)
    bool lencounter_sas_plpl = true;     // ENCOUNTER_CHECK_PLPL SORTSWEEP (swiftest_io.f90:2720-2744)
    bool lencounter_sas_pltp = true;
    bool lmtiny_pl = false;              // GMTINY set
    DP GMTINY = -1.0;
    bool lgr = false;                    // GR
    DP inv_c2 = 0.0;
};

// base_util_exit(FAILURE) stand-in (base/base_module.f90:589): a failed device call is fatal
struct fatal_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class cuda_context {
   public:
    explicit cuda_context(int device = 0)
    {
        const int rc = swcu_create(device, &h_);
        if (rc != SWCU_OK)
            throw fatal_error(rc == SWCU_ERR_NOGPU ? "swiftest_cuda: no sm_100 GPU (there is no CPU fallback)"
                                                   : "swiftest_cuda: swcu_create failed");
    }
    ~cuda_context()
    {
        if (h_) swcu_destroy(h_);
    }
    cuda_context(const cuda_context &) = delete;
    cuda_context &operator=(const cuda_context &) = delete;
    swcu_context *handle() const { return h_; }
    void check(int status, const char *where) const
    {
        if (status != SWCU_OK) throw fatal_error(std::string("swiftest_cuda: ") + where + ": " + swcu_last_error(h_));
    }

   private:
    swcu_context *h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// generic swiftest_kick_getacch_int_all (swiftest_module.f90:940-991): nplpl is I8B => flat, nplm is I4B => triangular,
// presence of radius => radius-checked variant
// ---------------------------------------------------------------------------------------------------------------------
// swiftest_kick_getacch_int_all_flat_rad_pl (kick.f90:69-115).  k_plpl == nullptr: the canonical flattened pairs.
inline void swiftest_kick_getacch_int_all(cuda_context &c, I4B npl, I8B nplpl, const I4B *k_plpl, const DP *r,
                                          const DP *Gmass, const DP *radius, DP *acc)
{
    c.check(swcu_kick_getacch_int_all_flat_pl(c.handle(), npl, nplpl, k_plpl, r, Gmass, radius, acc), "flat_rad_pl");
}
// swiftest_kick_getacch_int_all_flat_norad_pl (kick.f90:118-162)
inline void swiftest_kick_getacch_int_all(cuda_context &c, I4B npl, I8B nplpl, const I4B *k_plpl, const DP *r,
                                          const DP *Gmass, DP *acc)
{
    c.check(swcu_kick_getacch_int_all_flat_pl(c.handle(), npl, nplpl, k_plpl, r, Gmass, nullptr, acc), "flat_norad_pl");
}
// swiftest_kick_getacch_int_all_tri_rad_pl (kick.f90:165-271)
inline void swiftest_kick_getacch_int_all(cuda_context &c, I4B npl, I4B nplm, const DP *r, const DP *Gmass,
                                          const DP *radius, DP *acc)
{
    c.check(swcu_kick_getacch_int_all_tri_pl(c.handle(), npl, nplm, r, Gmass, radius, acc), "tri_rad_pl");
}
// swiftest_kick_getacch_int_all_tri_norad_pl (kick.f90:274-371)
inline void swiftest_kick_getacch_int_all(cuda_context &c, I4B npl, I4B nplm, const DP *r, const DP *Gmass, DP *acc)
{
    c.check(swcu_kick_getacch_int_all_tri_pl(c.handle(), npl, nplm, r, Gmass, nullptr, acc), "tri_norad_pl");
}
// swiftest_kick_getacch_int_all_tp (kick.f90:374-415)
inline void swiftest_kick_getacch_int_all_tp(cuda_context &c, I4B ntp, I4B npl, const DP *rtp, const DP *rpl,
                                             const DP *GMpl, const I4B *lmask, DP *acc)
{
    c.check(swcu_kick_getacch_int_all_tp(c.handle(), ntp, npl, rtp, rpl, GMpl, lmask, acc), "all_tp");
}

// swiftest_drift_all (drift.f90:60-108)
inline void swiftest_drift_all(cuda_context &c, const DP *mu, DP *x, DP *v, I4B n, const swiftest_parameters &param, DP dt,
                               const I4B *lmask, I4B *iflag)
{
    if (n == 0) return;
    c.check(swcu_drift_all(c.handle(), n, mu, x, v, dt, param.lgr ? 1 : 0, param.inv_c2, lmask, iflag), "drift_all");
}

// encounter_list (encounter/encounter_module.f90:23-47): the part the detection fills
struct encounter_list {
    I8B nenc = 0;
    std::vector<I4B> index1, index2;
    std::vector<I4B> lvdotr;  // Fortran logical as 0/1
    void resize(I8B n)
    {
        nenc = n;
        index1.resize((size_t)n);
        index2.resize((size_t)n);
        lvdotr.resize((size_t)n);
    }
};

namespace detail {
inline void fetch(cuda_context &c, I8B nenc, encounter_list &out)
{
    out.resize(nenc);  // the Fortran caller allocates its intent(out) arrays here, after learning nenc
    if (nenc > 0)
        c.check(swcu_encounter_fetch(c.handle(), nenc, out.index1.data(), out.index2.data(), out.lvdotr.data()),
                "encounter_fetch");
}
}  // namespace detail

// encounter_check_all_plpl (encounter_check.f90:14-39): SORTSWEEP or TRIANGULAR by param%lencounter_sas_plpl
inline void encounter_check_all_plpl(cuda_context &c, const swiftest_parameters &param, I4B npl, const DP *r, const DP *v,
                                     const DP *renc, DP dt, encounter_list &out)
{
    I8B nenc = 0;
    if (param.lencounter_sas_plpl)
        c.check(swcu_encounter_check_all_sort_and_sweep_plpl(c.handle(), npl, r, v, renc, dt, &nenc), "sas_plpl");
    else
        c.check(swcu_encounter_check_all_triangular_plpl(c.handle(), npl, r, v, renc, dt, &nenc), "tri_plpl");
    detail::fetch(c, nenc, out);
}
// encounter_check_all_plplm (encounter_check.f90:42-109): plpl on the fully interacting block + plm x plt, index2
// shifted by nplm, merged.  The sort-and-sweep form merges on the device; the triangular form merges here like :77-103
// (the reference orders by index1 only; lexicographic order is one of the orders it allows).
inline void encounter_check_all_plplm(cuda_context &c, const swiftest_parameters &param, I4B nplm, I4B nplt, const DP *rplm,
                                      const DP *vplm, const DP *rplt, const DP *vplt, const DP *rencm, const DP *renct,
                                      DP dt, encounter_list &out)
{
    I8B nenc = 0;
    if (param.lencounter_sas_plpl) {
        c.check(swcu_encounter_check_all_plplm(c.handle(), nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt, &nenc),
                "all_plplm");
        detail::fetch(c, nenc, out);
        return;
    }
    encounter_list a, b;
    c.check(swcu_encounter_check_all_triangular_plpl(c.handle(), nplm, rplm, vplm, rencm, dt, &nenc), "tri_plpl");
    detail::fetch(c, nenc, a);
    c.check(swcu_encounter_check_all_triangular_plplm(c.handle(), nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt, &nenc),
            "tri_plplm");
    detail::fetch(c, nenc, b);
    std::vector<std::pair<I4B, I4B>> all;
    all.reserve((size_t)(a.nenc + b.nenc));
    for (I8B k = 0; k < a.nenc; ++k) all.emplace_back(a.index1[k], a.index2[k]);
    for (I8B k = 0; k < b.nenc; ++k) all.emplace_back(b.index1[k], b.index2[k] + nplm);
    std::sort(all.begin(), all.end());
    out.resize((I8B)all.size());
    for (size_t k = 0; k < all.size(); ++k) {
        out.index1[k] = all[k].first;
        out.index2[k] = all[k].second;
        out.lvdotr[k] = 1;
    }
}
// encounter_check_all_pltp (encounter_check.f90:112-140)
inline void encounter_check_all_pltp(cuda_context &c, const swiftest_parameters &param, I4B npl, I4B ntp, const DP *rpl,
                                     const DP *vpl, const DP *rtp, const DP *vtp, const DP *renc, DP dt,
                                     encounter_list &out)
{
    I8B nenc = 0;
    if (param.lencounter_sas_pltp)
        c.check(swcu_encounter_check_all_sort_and_sweep_pltp(c.handle(), npl, ntp, rpl, vpl, rtp, vtp, renc, dt, &nenc),
                "sas_pltp");
    else
        c.check(swcu_encounter_check_all_triangular_pltp(c.handle(), npl, ntp, rpl, vpl, rtp, vtp, renc, dt, &nenc),
                "tri_pltp");
    detail::fetch(c, nenc, out);
}

// swiftest_util_get_potential_energy (generic of _flat / _triangular, swiftest_util.f90:1291-1394)
inline DP swiftest_util_get_potential_energy(cuda_context &c, I4B npl, const I4B *lmask, DP GMcb, const DP *Gmass,
                                             const DP *mass, const DP *rb)
{
    DP pe = 0.0;
    c.check(swcu_util_get_potential_energy(c.handle(), npl, lmask, GMcb, Gmass, mass, rb, &pe), "get_potential_energy");
    return pe;
}

// the double loop of swiftest_discard_pl_tp (swiftest_discard.f90:261-288): iplanet(i) = discarding planet or 0
inline I4B swiftest_discard_pl_tp(cuda_context &c, I4B ntp, I4B npl, const DP *rtp, const DP *vtp, const I4B *lactive,
                                  const DP *rpl, const DP *vpl, const DP *radius, DP dt, std::vector<I4B> &iplanet)
{
    iplanet.assign((size_t)ntp, 0);
    I4B nd = 0;
    c.check(swcu_discard_pl_tp(c.handle(), ntp, npl, rtp, vtp, lactive, rpl, vpl, radius, dt, iplanet.data(), &nd),
            "discard_pl_tp");
    return nd;
}

// ---------------------------------------------------------------------------------------------------------------------
// body types with the type-bound procedures of the hot path
// ---------------------------------------------------------------------------------------------------------------------
struct swiftest_cb {
    DP Gmass = 0.0;
};

struct swiftest_body {  // swiftest_module.f90:123-232
    I4B nbody = 0;
    std::vector<DP> rh, vh, vb, ah, mu;  // (3,n) except mu(n)
    std::vector<I4B> lmask, iflag;
    virtual ~swiftest_body() = default;
    virtual void setup(I4B n)
    {
        nbody = n;
        rh.assign(3 * (size_t)n, 0.0);
        vh = vb = ah = rh;
        mu.assign((size_t)n, 0.0);
        lmask.assign((size_t)n, 1);
        iflag.assign((size_t)n, 0);
    }
    // swiftest_drift_body (drift.f90:21-57): drifts rh, vh with self%mu; returns the number of bodies lost
    virtual int drift(cuda_context &c, const swiftest_parameters &param, DP dt)
    {
        std::fill(iflag.begin(), iflag.end(), 0);
        swiftest_drift_all(c, mu.data(), rh.data(), vh.data(), nbody, param, dt, lmask.data(), iflag.data());
        int lost = 0;
        for (I4B f : iflag) lost += (f != 0);  // caller sets status = DISCARDED_DRIFTERR (drift.f90:43-50)
        return lost;
    }
};

struct swiftest_pl : swiftest_body {  // swiftest_module.f90:234-294
    std::vector<DP> Gmass, radius, rhill, renc;
    I4B nplm = 0;
    I8B nplpl = 0, nplplm = 0;
    void setup(I4B n) override
    {
        swiftest_body::setup(n);
        Gmass.assign((size_t)n, 0.0);
        radius = rhill = renc = Gmass;
        nplm = n;
    }
    // swiftest_util_flatten_eucl_plpl + symba_util_flatten_eucl_plpl (swiftest_util.f90:1090-1130, symba_util.f90:175-208):
    // pair counts only -- the k_plpl table itself is never built on this path
    void flatten(const swiftest_parameters &param)
    {
        const I8B n = nbody;
        I8B m = n;
        if (param.lmtiny_pl) {
            m = 0;
            for (DP g : Gmass) m += (g >= param.GMTINY);  // bodies are kept sorted by mass, descending (:1709)
        }
        nplm = (I4B)m;
        nplpl = n * (n - 1) / 2;
        nplplm = m * n - m * (m + 1) / 2;
    }
    // swiftest_kick_getacch_int_pl (kick.f90:12-43)
    virtual void accel_int(cuda_context &c, const swiftest_parameters &param)
    {
        if (nbody == 0) return;
        if (param.lflatten_interactions) {
            if (param.lclose)
                swiftest_kick_getacch_int_all(c, nbody, nplpl, nullptr, rh.data(), Gmass.data(), radius.data(), ah.data());
            else
                swiftest_kick_getacch_int_all(c, nbody, nplpl, nullptr, rh.data(), Gmass.data(), ah.data());
        } else {
            if (param.lclose)
                swiftest_kick_getacch_int_all(c, nbody, nbody, rh.data(), Gmass.data(), radius.data(), ah.data());
            else
                swiftest_kick_getacch_int_all(c, nbody, nbody, rh.data(), Gmass.data(), ah.data());
        }
    }
    // helio_drift_body (helio/helio_drift.f90:14-54): drift rh, vb with mu = cb%Gmass for every body
    int helio_drift(cuda_context &c, const swiftest_cb &cb, const swiftest_parameters &param, DP dt)
    {
        std::vector<DP> mucb((size_t)nbody, cb.Gmass);
        std::fill(iflag.begin(), iflag.end(), 0);
        swiftest_drift_all(c, mucb.data(), rh.data(), vb.data(), nbody, param, dt, lmask.data(), iflag.data());
        int lost = 0;
        for (I4B f : iflag) lost += (f != 0);
        return lost;
    }
};

struct swiftest_tp : swiftest_body {  // swiftest_module.f90:296-329
    // swiftest_kick_getacch_int_tp (kick.f90:46-66)
    void accel_int(cuda_context &c, const swiftest_parameters &, const DP *GMpl, const DP *rhp, I4B npl)
    {
        if (nbody == 0 || npl == 0) return;
        swiftest_kick_getacch_int_all_tp(c, nbody, npl, rh.data(), rhp, GMpl, lmask.data(), ah.data());
    }
};

struct symba_pl : swiftest_pl {  // symba/symba_module.f90:30-52
    static constexpr DP RHSCALE = 6.5, RSHELL = 0.48075;  // symba_module.f90:22-23
    // symba_kick_getacch_int_pl (symba_kick.f90:14-33): always the radius-checked variants
    void accel_int(cuda_context &c, const swiftest_parameters &param) override
    {
        if (nbody == 0) return;
        if (param.lflatten_interactions)
            swiftest_kick_getacch_int_all(c, nbody, nplplm, nullptr, rh.data(), Gmass.data(), radius.data(), ah.data());
        else
            swiftest_kick_getacch_int_all(c, nbody, nplm, rh.data(), Gmass.data(), radius.data(), ah.data());
    }
    // symba_util_set_renc (symba_util.f90:245-267)
    void set_renc(I4B scale)
    {
        DP rshell_irec = 1.0;
        for (I4B i = 1; i <= scale; ++i) rshell_irec = rshell_irec * RSHELL;
        for (I4B i = 0; i < nbody; ++i) renc[(size_t)i] = rhill[(size_t)i] * RHSCALE * rshell_irec;
    }
    // the hot part of symba_kick_getacch_pl (symba_kick.f90:36-76): interaction accelerations of all bodies, then
    // the pairs of the encounter list once more with flat_rad, subtracted
    void kick_getacch(cuda_context &c, const swiftest_parameters &param, const encounter_list &plpl_encounter)
    {
        if (nbody == 0) return;
        accel_int(c, param);
        if (plpl_encounter.nenc > 0)
            c.check(swcu_symba_kick_subtract_encounters(c.handle(), nbody, plpl_encounter.nenc, plpl_encounter.index1.data(),
                                                        plpl_encounter.index2.data(), rh.data(), Gmass.data(), radius.data(),
                                                        ah.data()),
                    "symba_kick_getacch_pl");
    }
    // the detection part of symba_encounter_check_pl (symba_encounter_check.f90:14-52); returns lany_encounter
    bool encounter_check(cuda_context &c, const swiftest_parameters &param, DP dt, I4B irec, encounter_list &plpl_encounter)
    {
        if (nbody == 0) return false;
        const I4B npl = nbody, nplt = npl - nplm;
        set_renc(irec);
        if (nplt == 0)
            encounter_check_all_plpl(c, param, npl, rh.data(), vb.data(), renc.data(), dt, plpl_encounter);
        else
            encounter_check_all_plplm(c, param, nplm, nplt, rh.data(), vb.data(), rh.data() + 3 * (size_t)nplm,
                                      vb.data() + 3 * (size_t)nplm, renc.data(), renc.data() + nplm, dt, plpl_encounter);
        return plpl_encounter.nenc > 0;
    }
};

struct symba_tp : swiftest_tp {  // symba/symba_module.f90:54-66
    // the detection part of symba_encounter_check_tp (symba_encounter_check.f90:238-265)
    bool encounter_check(cuda_context &c, const swiftest_parameters &param, symba_pl &pl, DP dt, I4B irec,
                         encounter_list &pltp_encounter)
    {
        if (nbody == 0) return false;
        pl.set_renc(irec);
        encounter_check_all_pltp(c, param, pl.nbody, nbody, pl.rh.data(), pl.vb.data(), rh.data(), vb.data(),
                                 pl.renc.data(), dt, pltp_encounter);
        return pltp_encounter.nenc > 0;
    }
};

}  // namespace swiftest
