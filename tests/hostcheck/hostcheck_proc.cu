// TEST INFRASTRUCTURE: runs a generated process's Proc::matrix() on the CPU.
// The device functions are __host__ __device__, so the very code the GPU kernel executes can be
// checked against the oracle in the build container (no GPU there).  Never linked into the product.
#include <vector>

#include MF_PROC_SOURCE

extern "C" int hostcheck_smatrix(const double* p, long long nevt, const double* par, const double* coup,
                                 long long coup_stride, double sqh, int only_comb, double* out) {
  for (long long ev = 0; ev < nevt; ++ev) {
    double m[Proc::NEXT][4];
    for (int i = 0; i < Proc::NEXT; ++i)
      for (int k = 0; k < 4; ++k) m[i][k] = p[(ev * Proc::NEXT + i) * 4 + k];
    cxd c[Proc::NCOUP > 0 ? Proc::NCOUP : 1];
    for (int j = 0; j < Proc::NCOUP; ++j) {
      const long long o = coup_stride ? ((long long)j * nevt + ev) : j;
      c[j] = mk(coup[2 * o], coup[2 * o + 1]);
    }
    if (only_comb >= 0)
      out[ev] = Proc::matrix(m, only_comb, par, c, sqh);
    else
      out[ev] = mf::smatrix_event<Proc>(m, par, c, sqh);
  }
  return 0;
}

// The helicity-parallel flavour: its phases are separated by block barriers only, so running every
// "thread" of a phase in turn on the CPU reproduces the kernel's data flow exactly.
extern "C" int hostcheck_smatrix_hp(const double* p, long long nevt, const double* par, const double* coup,
                                    long long coup_stride, double sqh, int only_comb, double* out) {
  constexpr int E = Proc::HP_E, NH = Proc::NCOMB, NCG = Proc::HP_NCG, EVS = Proc::HP_EVSTRIDE;
  std::vector<cxd> evarea((size_t)EVS * E);
  std::vector<double> mom(E * Proc::NEXT * 4);
  std::vector<cxd> cp(E * (Proc::NCOUP > 0 ? Proc::NCOUP : 1));
  std::vector<cxd> ftab((size_t)E * mf::SLU_NF, mk(0.0, 0.0));
  // one phase of units: the table-driven routine, or the straight-line units warp by warp, lane by lane
  auto units = [&](int slu_phase, int begin, int n) {
    if (Proc::HP_SLU) {
      for (int w = 0; w < Proc::HP_THREADS / 32; ++w)
        for (int l = 0; l < 32; ++l) mf::slu_units<Proc>(slu_phase, w, l, ftab.data(), par, evarea.data());
    } else {
      mf::hp_units<Proc, Proc::HP_SPLIT>(begin, n, 0, 1, par, cp.data(), evarea.data());
    }
  };
  std::vector<unsigned char> vtab((1 << Proc::NEXT) * NH);
  for (int i = 0; i < (1 << Proc::NEXT) * NH; ++i) mf::hp_fill_vtab<Proc>(i, vtab.data());
  int only_h = -1;
  if (only_comb >= 0) {
    only_h = 0;
    for (int j = 0; j < Proc::NEXT; ++j) only_h |= ((Proc::hel(only_comb, j) + 1) >> 1) << j;
  }
  for (long long ev0 = 0; ev0 < nevt; ev0 += E) {
    const int nev = (int)((nevt - ev0) < E ? (nevt - ev0) : E);
    for (int e = 0; e < E; ++e) {
      const long long ev = ev0 + (e < nev ? e : 0);
      for (int r = 0; r < Proc::NEXT * 4; ++r) mom[e * Proc::NEXT * 4 + r] = p[ev * Proc::NEXT * 4 + r];
      for (int j = 0; j < Proc::NCOUP; ++j) {
        const long long o = coup_stride ? ((long long)j * nevt + ev) : j;
        cp[e * Proc::NCOUP + j] = mk(coup[2 * o], coup[2 * o + 1]);
        mf::hp_fill_ftab(cp[e * Proc::NCOUP + j], ftab.data() + e * mf::SLU_NF + 4 * j);
      }
    }
    for (int it = 0; it < Proc::NEXT * E * 2; ++it) mf::hp_externals<Proc>(it, E, mom.data(), par, sqh, evarea.data());
    for (int L = 2; L <= Proc::HP_MAXLEVEL; ++L) {
      units(L - 2, Proc::level_begin(L), Proc::level_begin(L + 1) - Proc::level_begin(L));
    }
    std::vector<double> me_h((size_t)E * NH, 0.0);
    if (Proc::HP_UNROLL) {
      for (int e = 0; e < nev; ++e)
        for (int h = 0; h < NH; ++h)
          me_h[e * NH + h] = Proc::hp_amps_unrolled(evarea.data() + e * EVS, vtab.data(), h, cp.data() + e * Proc::NCOUP);
    } else {
      // "thread" t = (e * NCG + cg) * NHP + h owns HP_NJ JAMPs; one helicity pass after the other
      constexpr int NJ = Proc::HP_NJ, NHP = Proc::HP_NHP, T = E * NCG * NHP;
      for (int pass = 0; pass < Proc::HP_NPASS; ++pass) {
        std::vector<cxd> J((size_t)T * NJ, mk(0.0, 0.0));
        for (int bi = 0; bi < Proc::HP_NBATCH; ++bi) {
          const mf::HpBatch bt = Proc::batch(pass * Proc::HP_NBATCH + bi);
          units(Proc::HP_MAXLEVEL - 1 + pass * Proc::HP_NBATCH + bi, bt.unit_begin, bt.unit_end - bt.unit_begin);
          // the tiles of the batch in table order
          for (int ee = 0; ee < E; ++ee)
            for (int ti = bt.tile_begin; ti < bt.tile_end; ++ti) {
              cxd* a_e = evarea.data() + ee * EVS;
              mf::hp_mma_tile_host<Proc>(Proc::tile(ti), a_e, a_e + Proc::HP_WFSIZE + Proc::HP_SCRATCH);
            }
          for (int t = 0; t < T; ++t) {
            const int e = t / (NCG * NHP), cg = (t / NHP) % NCG, h = t % NHP;
            cxd(&Jt)[NJ] = *reinterpret_cast<cxd(*)[NJ]>(&J[(size_t)t * NJ]);
            Proc::jamp_batch(bi, cg, evarea.data() + e * EVS + Proc::HP_WFSIZE + Proc::HP_SCRATCH + mf::hp_abuf_pos(h), Jt);
          }
        }
        if (Proc::HP_COLOUR == 1)
          for (int t = 0; t < T; ++t) {
            const int e = t / (NCG * NHP), cg = (t / NHP) % NCG, h = t % NHP;
            for (int j = 0; j < NJ; ++j)
              if (cg * NJ + j < Proc::NCOLOR) evarea[(size_t)e * EVS + Proc::HP_WFSIZE + h + (cg * NJ + j) * NHP] = J[(size_t)t * NJ + j];
          }
        if (Proc::HP_COLOUR >= 2) {
          constexpr int NCP = Proc::HP_NCP, PL = Proc::HP_PLANE;
          for (int e = 0; e < E; ++e) {
            double* planes = reinterpret_cast<double*>(evarea.data() + (size_t)e * EVS + Proc::HP_WFSIZE);
            for (int i = 0; i < 2 * NCP * PL; ++i) planes[i] = 0.0;
          }
          for (int t = 0; t < T; ++t) {
            const int e = t / (NCG * NHP), cg = (t / NHP) % NCG, h = t % NHP;
            double* planes = reinterpret_cast<double*>(evarea.data() + (size_t)e * EVS + Proc::HP_WFSIZE);
            for (int j = 0; j < NJ; ++j)
              if (cg * NJ + j < Proc::NCOLOR) {
                planes[(cg * NJ + j) * PL + h] = J[(size_t)t * NJ + j].re;
                planes[(NCP + cg * NJ + j) * PL + h] = J[(size_t)t * NJ + j].im;
              }
          }
        }
        for (int t = 0; t < T; ++t) {
          const int e = t / (NCG * NHP), cg = (t / NHP) % NCG, h = t % NHP;
          cxd(&Jt)[NJ] = *reinterpret_cast<cxd(*)[NJ]>(&J[(size_t)t * NJ]);
          double m;
          if (Proc::HP_COLOUR >= 2)  // both table-driven flavours compute sum_a J_a (cfsym J)_a
            m = mf::hp_colour_loop<Proc>(reinterpret_cast<const double*>(evarea.data() + (size_t)e * EVS + Proc::HP_WFSIZE), h, cg, Proc::cfsym());
          else
            m = Proc::colour_sum(cg, Jt, evarea.data() + e * EVS + Proc::HP_WFSIZE + h);
          me_h[e * NH + pass * NHP + h] += m;
        }
      }
    }
    for (int e = 0; e < nev; ++e) {
      double acc = 0.0;
      for (int h = 0; h < NH; ++h)
        if (only_h < 0 || h == only_h) acc += me_h[e * NH + h];
      out[ev0 + e] = only_h >= 0 ? acc : acc / Proc::DENOM;
    }
  }
  return 0;
}
