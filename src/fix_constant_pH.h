/* -*- c++ -*- ----------------------------------------------------------
   fix constant_pH -- B200 drop-in for MahdiTavakol/Constant_pH fix_constant_pH.h

   Same style name, same constructor arguments and the same hook surface as the
   reference class (reference fix_constant_pH.h:29-59); the per-timestep work is done by
   libcph_b200.so (include/cph_b200.h), hand-written sm_100a CUDA, instead of host loops
   over per-atom arrays.  Hooks the reference declares (h:31-40) keep their signatures;
   hooks north_star adds (initial_integrate, final_integrate, write_restart, restart,
   pack/unpack_reverse_comm, which the reference defines but never declares, cpp:287-308)
   are marked below.  Private state is this implementation's own; only the public surface
   is dictated by the reference.
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
  // ---- the reference's public surface (h:31-40) ------------------------------------------------
  FixConstantPH(class LAMMPS *, int, char **);         // cpp:33-56
  ~FixConstantPH() override;
  int setmask() override;                               // declared, never defined in the reference
  void init() override;                                 // cpp:83-105
  void setup(int) override;                             // declared, never defined
  void post_force(int) override;                        // cpp:67-79
  double compute_scalar() override;                     // declared, never defined: H_lambda of cpp:114
  double compute_vector(int) override;                  // declared, never defined
  double memory_usage() override;                       // cpp:314-318 (defined on the wrong class there)
  void init_list(int, class NeighList *) override;      // declared, never defined: the library owns its list

  // ---- hooks north_star asks for, absent from the reference ------------------------------------
  void initial_integrate(int) override;
  void final_integrate() override;
  void post_neighbor() override;
  void write_restart(FILE *) override;
  void restart(char *) override;
  int pack_reverse_comm(int, int, double *) override;           // body at cpp:287-295, undeclared there
  void unpack_reverse_comm(int, int *, double *) override;      // body at cpp:299-308, undeclared there

 private:
  // positional arguments arg[3..8] (cpp:37-49)
  struct Args {
    int hyd_group, wat_group;      // group ids of arg[4], arg[5]
    int hyd_bit, wat_bit;          // their bitmasks
    double pK, pH, temperature;
  } in;

  // Donnini-2016 bias constants loaded in init() (cpp:86-96)
  struct Bias {
    double w, s, h, k, a, b, r, m, d;
    double mass;                   // m_lambda
  } bias;

  // keyword-selected behaviour (the reference's keyword loop, cpp:51-54, is empty)
  struct Options {
    int dudl, integrator, fscale, bias_form, buffer, theta, excluded_drop;
    double thermostat_period;
    double lambda_start;
    char *site_file;
    double bias_user[9];           // w s h k a b r m d given as keywords (bias_w ... bias_d); NaN = table value
    int ewald_kmax[3];             // keyword ewald: wave-vector range of the device-side Ewald sum (0 = off)
  } opt;

  // per-site table read from the site file (north_star multi-site; none in the reference)
  struct Sites {
    int nsites, natoms;
    double *pK, *lambda0, *qA, *qB;
    int *tag, *site;
    int *typeB;                    // optional fifth column: atom type in state B (LJ end states), 0 = none
    int lj_states;
  } tab;

  cph_handle *cph;                 // the device side
  double part[2];                  // HA, HB of the last reduction (cpp:276-277)
  double lambda_cached;            // lambda of the reference's single site, for the host-side rescale
  double *host_energy;             // the reference's H_atom (h:51): host-tallied energy sources
  int host_energy_cap;             // its length (the reference's nmax)
  double *force_out;               // pair forces of the last pass (only fetched for the host-part rescale)
  int force_cap;
  double *charge_buf;              // q(lambda) on its way into atom->q
  int charge_cap;
  bool sites_on_device;            // site table + initial lambda sent (once, not at every run)
  bool rank_group_ready;           // NCCL rank group built over `world`
  double *pending_restart;         // restart record received before init()
  int pending_n;
  bool resend_atoms;

  // the reference's private helpers (h:53-58) that still have host-side work; calculate_df, calculate_dU,
  // integrate_lambda and modify_water (h:54-58) are kernels of the library (sites.cu)
  void compute_Hs();               // cpp:177-280: host-tallied sources only, the pair part is on the device
  void set_force();                // cpp:149-171
  void kspace_site_derivative(const double *ekspace);   // cpp:241-244 seen by the charge derivative

  void require(int rc, const char *what);
  void run_device_step(bool setup_only);
  bool host_forces_present() const;
  void scale_host_part();
  void pull_charges();
  void load_site_table(const char *path);
  void upload_atoms();
  void scale_hydrogen_forces(double factor);
};

}    // namespace LAMMPS_NS

#endif
#endif
