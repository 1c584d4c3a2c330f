/* -*- c++ -*- ----------------------------------------------------------
   fix constant_pH -- B200 drop-in for MahdiTavakol/Constant_pH fix_constant_pH.h

   Same style name, same constructor arguments and the same hook surface as the
   reference class (reference fix_constant_pH.h:29-59); the per-timestep work is done by
   libcph_b200.so (include/cph_b200.h), hand-written sm_100a CUDA, instead of host loops
   over per-atom arrays.  Hooks the reference declares (h:31-40) keep their signatures;
   hooks north_star adds (initial_integrate, final_integrate, write_restart, restart,
   pack/unpack_reverse_comm, which the reference defines but never declares, cpp:287-308)
   are marked below.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(constant_pH,FixConstantPH);
// clang-format on
#else

#ifndef LMP_FIX_CONSTANTPH_H
#define LMP_FIX_CONSTANTPH_H

#include "fix.h"

struct cph_handle;

namespace LAMMPS_NS {

class FixConstantPH : public Fix {
 public:
  FixConstantPH(class LAMMPS *, int, char **);         // reference h:31, cpp:33-56
  ~FixConstantPH() override;                            // h:32
  int setmask() override;                               // h:33 (never defined in the reference)
  void init() override;                                 // h:34, cpp:83-105
  void setup(int) override;                             // h:35 (never defined)
  void post_force(int) override;                        // h:36, cpp:67-79
  double compute_scalar() override;                     // h:37 (never defined): H_lambda of cpp:114
  double compute_vector(int) override;                  // h:38 (never defined)
  double memory_usage() override;                       // h:39, cpp:314-318 (defined on the wrong class there)
  void init_list(int, class NeighList *) override;      // h:40 (never defined): the library owns its list
  // north_star hooks absent from the reference
  void initial_integrate(int) override;
  void final_integrate() override;
  void post_neighbor() override;
  void write_restart(FILE *) override;
  void restart(char *) override;
  int pack_reverse_comm(int, int, double *) override;           // cpp:287-295
  void unpack_reverse_comm(int, int *, double *) override;      // cpp:299-308

 private:
  // Input variables for constant values (reference h:44-51)
  int igroupH, igroupW;
  int groupHbit, groupWbit;
  double pK, pH, T;
  double a, b, s, m, w, r, d, h, k;      // h and k are used at cpp:88-89 but undeclared in the reference
  double m_lambda;
  double t_lambda_period;                // Nose-Hoover period of the lambda thermostat (0 = off)
  double HA, HB;
  int nmax;
  double *H_atom;                        // kept for interface parity; energies live on the device

  // additions
  cph_handle *cph;
  int dudl_mode, integrator_mode, fscale_mode, bias_mode, water_buffer, coord_theta;
  char *sitefile;
  int nsites, ntitr;
  double *site_pK, *site_lambda0, *titr_qA, *titr_qB;
  int *titr_tag, *titr_site;
  double *restart_buf;
  int restart_n;
  bool atoms_sent;
  double *xbuf, *fbuf;
  int bufmax;
  double lambda_host;                    // lambda of the single reference site, for the host-side rescale

  // reference helpers (h:53-58); the arithmetic now runs in the library
  void integrate_lambda();               // cpp:109-117  -> cph_integrate_lambda
  void compute_Hs();                     // cpp:177-280  -> cph_pair_pass + cph_site_reduce
  void calculate_df();                   // cpp:120-124  -> lambda integrator kernel
  void calculate_dU();                   // cpp:128-145  -> lambda integrator kernel
  void set_force();                      // cpp:149-171
  void modify_water();                   // h:58, never defined nor called in the reference

  void check(int rc, const char *what);
  void read_sites(const char *path);
  void send_atoms();
};

}    // namespace LAMMPS_NS

#endif
#endif
