// lammps_shim.h -- the slice of the LAMMPS class surface that src/fix_constant_pH.{h,cpp} touches.
//
// LAMMPS is not available in this environment, so the drop-in fix is compiled
// against these declarations and driven by src/harness.cpp.  Member names, types and call
// signatures follow upstream LAMMPS (stable 2Aug2023 conventions, from memory -- no source to
// cite) so that fix_constant_pH.cpp compiles unchanged inside a real LAMMPS tree, where the
// real "fix.h", "atom.h", ... are found instead of the forwarding headers next to this file.
// Nothing here is product code; the product is the fix source and libcph_b200.so.
#ifndef LMP_SHIM_H
#define LMP_SHIM_H

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <vector>

#define FLERR __FILE__, __LINE__

// ---- the three MPI names the fix uses (upstream: <mpi.h> through lmptype.h).  The shim's "world" is a set of
// harness processes that share a scratch directory: MPI_Bcast is a file the root writes and the others wait for.
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_BYTE 1
#define MPI_INT 4
struct ShimWorld {
  int rank = 0, size = 1, seq = 0;
  std::string dir;
  static ShimWorld &get() {
    static ShimWorld w;
    static bool init = false;
    if (!init) {
      init = true;
      if (const char *e = getenv("CPH_SHIM_RANK")) w.rank = atoi(e);
      if (const char *e = getenv("CPH_SHIM_NRANKS")) w.size = atoi(e);
      if (const char *e = getenv("CPH_SHIM_DIR")) w.dir = e;
    }
    return w;
  }
};
inline int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm) {
  ShimWorld &w = ShimWorld::get();
  if (w.size <= 1) return 0;
  const size_t bytes = (size_t)count * (size_t)type;
  const std::string path = w.dir + "/bcast_" + std::to_string(w.seq++);
  if (w.rank == root) {
    const std::string tmp = path + ".tmp";
    FILE *fp = fopen(tmp.c_str(), "wb");
    if (!fp || fwrite(buf, 1, bytes, fp) != bytes) { fprintf(stderr, "shim MPI_Bcast: cannot write %s\n", tmp.c_str()); exit(4); }
    fclose(fp);
    rename(tmp.c_str(), path.c_str());
  } else {
    for (int tries = 0;; tries++) {
      FILE *fp = fopen(path.c_str(), "rb");
      if (fp) {
        const size_t got = fread(buf, 1, bytes, fp);
        fclose(fp);
        if (got == bytes) break;
      }
      if (tries > 60000) { fprintf(stderr, "shim MPI_Bcast: timed out on %s\n", path.c_str()); exit(4); }
      struct timespec ts = {0, 1000000};
      nanosleep(&ts, nullptr);
    }
  }
  return 0;
}

namespace LAMMPS_NS {

typedef int tagint;      // -DLAMMPS_SMALLBIG
typedef int64_t bigint;

struct LammpsAbort : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class LAMMPS;
class NeighList;

class Error {
 public:
  // upstream: [[noreturn]] void all(const std::string &file, int line, const std::string &fmt, args...)
  [[noreturn]] void all(const char *file, int line, const std::string &msg) {
    throw LammpsAbort(std::string("ERROR: ") + msg + " (" + file + ":" + std::to_string(line) + ")");
  }
  template <typename... Args>
  [[noreturn]] void all(const char *file, int line, const std::string &fmt, Args... args) {
    // "{}" placeholders as in upstream's fmtlib usage
    std::string out;
    std::vector<std::string> a = {tostr(args)...};
    size_t k = 0;
    for (size_t p = 0; p < fmt.size(); p++) {
      if (fmt[p] == '{' && p + 1 < fmt.size() && fmt[p + 1] == '}' && k < a.size()) { out += a[k++]; p++; }
      else out += fmt[p];
    }
    all(file, line, out);
  }
  void warning(const char *, int, const std::string &msg) { fprintf(stderr, "WARNING: %s\n", msg.c_str()); }

 private:
  static std::string tostr(const std::string &s) { return s; }
  static std::string tostr(const char *s) { return s; }
  template <typename T> static std::string tostr(T v) { return std::to_string(v); }
};

class Memory {
 public:
  template <typename T> T *create(T *&array, int n, const char *) { array = (T *)malloc(sizeof(T) * (n > 0 ? n : 1)); return array; }
  template <typename T> void destroy(T *&array) { free(array); array = nullptr; }
};

class Atom {
 public:
  int nlocal = 0, nghost = 0, nmax = 0;
  bigint natoms = 0;
  int ntypes = 0;
  double **x = nullptr, **f = nullptr;
  double *q = nullptr;
  int *type = nullptr, *mask = nullptr;
  tagint *tag = nullptr, *molecule = nullptr;
  int **nspecial = nullptr;
  tagint **special = nullptr;
  int maxspecial = 0;
  int q_flag = 1, molecule_flag = 1;
  enum { MAP_NONE = 0, MAP_ARRAY = 1, MAP_HASH = 2, MAP_YES = 3 };
  int map_style = MAP_ARRAY;
  std::vector<int> map_array;      // harness: tag -> local index (owned copy), -1 when not on this rank
  int map(tagint t) const { return (t >= 0 && (size_t)t < map_array.size()) ? map_array[t] : -1; }
};

class Group {
 public:
  std::vector<std::string> names{"all"};
  std::vector<int> bitmask_v{1};
  Atom *atom = nullptr;
  int find(const std::string &name) {
    for (size_t i = 0; i < names.size(); i++) if (names[i] == name) return (int)i;
    return -1;
  }
  int *bitmask = nullptr;          // upstream: int *bitmask (indexed by group id)
  bigint count_override = 0;       // harness, several ranks: upstream's count() is a sum over all ranks
  bigint count(int igroup) {
    if (count_override > 0 && igroup > 0) return count_override;
    bigint n = 0;
    for (int i = 0; i < atom->nlocal; i++) if (atom->mask[i] & bitmask[igroup]) n++;
    return n;
  }
  int add(const std::string &name, int bit) {
    names.push_back(name); bitmask_v.push_back(bit); bitmask = bitmask_v.data();
    return (int)names.size() - 1;
  }
};

class Pair {
 public:
  int compute_flag = 1;            // pair_modify compute yes/no
  double *eatom = nullptr;
  std::string style;
  // coefficients exposed the way pair_lj_cut_coul_cut.cpp's extract() does
  double **epsilon = nullptr, **sigma = nullptr;
  double cut_coul = 0, cut_lj_global = 0, alpha = 0;
  virtual ~Pair() {}
  virtual void *extract(const char *name, int &dim) {
    dim = 2;
    if (!strcmp(name, "epsilon")) return (void *)epsilon;
    if (!strcmp(name, "sigma")) return (void *)sigma;
    dim = 0;
    if (!strcmp(name, "cut_coul")) return (void *)&cut_coul;
    if (!strcmp(name, "cut_lj")) return (void *)&cut_lj_global;     // not upstream: see INTEGRATION.md
    if (!strcmp(name, "alpha")) return (void *)&alpha;              // not upstream: see INTEGRATION.md
    return nullptr;
  }
};

// bonded styles and KSpace: only what compute_Hs touches (cpp:221-244)
class EnergyStyle {
 public:
  double *eatom = nullptr;
  int compute_flag = 1;
  int tip4pflag = 0;
  double g_ewald = 0.0;            // KSpace only (public upstream): the Ewald splitting parameter
};
typedef EnergyStyle Bond;
typedef EnergyStyle Angle;
typedef EnergyStyle Dihedral;
typedef EnergyStyle Improper;
typedef EnergyStyle KSpace;

class Force {
 public:
  double boltz = 0.0019872067, qqrd2e = 332.06371, ftm2v = 1.0 / 48.88821291 / 48.88821291;
  double special_lj[4] = {1, 0, 0, 0}, special_coul[4] = {1, 0, 0, 0};
  int newton = 1, newton_pair = 1, newton_bond = 1;
  Pair *pair = nullptr;
  char *pair_style = nullptr;
  Bond *bond = nullptr;
  Angle *angle = nullptr;
  Dihedral *dihedral = nullptr;
  Improper *improper = nullptr;
  KSpace *kspace = nullptr;
  Pair *pair_match(const std::string &word, int exact, int = 0) {
    if (!pair) return nullptr;
    if (exact ? pair->style == word : pair->style.find(word) != std::string::npos) return pair;
    return nullptr;
  }
};

class Update {
 public:
  bigint ntimestep = 0;
  double dt = 1.0;
  bigint eflag_atom = 0;
};

class Domain {
 public:
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {0, 0, 0}, sublo[3] = {0, 0, 0}, subhi[3] = {0, 0, 0};
  int periodicity[3] = {1, 1, 1};
  int triclinic = 0;
};

class Comm {
 public:
  int me = 0, nprocs = 1;
  int procgrid[3] = {1, 1, 1}, myloc[3] = {0, 0, 0};
  // Ghosts of the harness are periodic images of the rank's own atoms (upstream: the swaps with sendproc == me).
  // ghost g = atom nlocal+g is an image of owned atom ghost_owner[g]; reverse_comm folds what the fix tallied on
  // the ghosts back onto their owners through the fix's own pack/unpack_reverse_comm, as Comm::reverse_comm(Fix *)
  // does upstream (defined below, after class Fix).
  std::vector<int> ghost_owner;
  int first_ghost = 0;
  inline void reverse_comm(class Fix *);
  void forward_comm(class Fix *) {}
};

class Neighbor {
 public:
  double skin = 2.0;
  int ago = 0;
};

class Modify {
 public:
  int n_energy_atom = 0;
};

class Universe {
 public:
  int me = 0;
};

class LAMMPS {
 public:
  Memory *memory;
  Error *error;
  Universe *universe;
  Atom *atom;
  Update *update;
  Neighbor *neighbor;
  Comm *comm;
  Domain *domain;
  Force *force;
  Modify *modify;
  Group *group;
  int world = 0;   // MPI_Comm in upstream
  LAMMPS() {
    memory = new Memory; error = new Error; universe = new Universe; atom = new Atom; update = new Update;
    neighbor = new Neighbor; comm = new Comm; domain = new Domain; force = new Force; modify = new Modify;
    group = new Group; group->atom = atom; group->bitmask = group->bitmask_v.data();
  }
  ~LAMMPS() {
    delete memory; delete error; delete universe; delete atom; delete update; delete neighbor; delete comm;
    delete domain; delete force; delete modify; delete group;
  }
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), universe(ptr->universe), atom(ptr->atom),
        update(ptr->update), neighbor(ptr->neighbor), comm(ptr->comm), domain(ptr->domain), force(ptr->force),
        modify(ptr->modify), group(ptr->group), world(ptr->world) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Universe *&universe;
  Atom *&atom;
  Update *&update;
  Neighbor *&neighbor;
  Comm *&comm;
  Domain *&domain;
  Force *&force;
  Modify *&modify;
  Group *&group;
  int &world;
};

namespace FixConst {
enum {
  INITIAL_INTEGRATE = 1 << 0, POST_INTEGRATE = 1 << 1, PRE_EXCHANGE = 1 << 2, PRE_NEIGHBOR = 1 << 3,
  POST_NEIGHBOR = 1 << 4, PRE_FORCE = 1 << 5, PRE_REVERSE = 1 << 6, POST_FORCE = 1 << 7,
  FINAL_INTEGRATE = 1 << 8, END_OF_STEP = 1 << 9
};
}

class Fix : protected Pointers {
 public:
  char *id = nullptr, *style = nullptr;
  int igroup = 0, groupbit = 1;
  int nevery = 1;
  int scalar_flag = 0, vector_flag = 0, size_vector = 0, global_freq = 0, extscalar = 0, extvector = 0;
  int restart_global = 0, comm_forward = 0, comm_reverse = 0, energy_global_flag = 0, thermo_energy = 0;
  int time_integrate = 0, dynamic_group_allow = 0, virial_global_flag = 0;

  Fix(LAMMPS *lmp, int narg, char **arg) : Pointers(lmp) {
    if (narg >= 3) { id = strdup(arg[0]); style = strdup(arg[2]); }
  }
  ~Fix() override { free(id); free(style); }
  virtual int setmask() = 0;
  virtual void init() {}
  virtual void init_list(int, NeighList *) {}
  virtual void setup(int) {}
  virtual void initial_integrate(int) {}
  virtual void post_neighbor() {}
  virtual void post_force(int) {}
  virtual void final_integrate() {}
  virtual void write_restart(FILE *) {}
  virtual void restart(char *) {}
  virtual int pack_forward_comm(int, int *, double *, int, int *) { return 0; }
  virtual void unpack_forward_comm(int, int, double *) {}
  virtual int pack_reverse_comm(int, int, double *) { return 0; }
  virtual void unpack_reverse_comm(int, int *, double *) {}
  virtual double compute_scalar() { return 0.0; }
  virtual double compute_vector(int) { return 0.0; }
  virtual double memory_usage() { return 0.0; }
};

inline void Comm::reverse_comm(Fix *fix) {
  const int n = (int)ghost_owner.size();
  if (n == 0) return;
  std::vector<double> buf((size_t)n * (fix->comm_reverse > 0 ? fix->comm_reverse : 1));
  fix->pack_reverse_comm(n, first_ghost, buf.data());
  fix->unpack_reverse_comm(n, ghost_owner.data(), buf.data());
}

namespace utils {
inline void missing_cmd_args(const std::string &file, int line, const std::string &cmd, Error *error) {
  error->all(file.c_str(), line, "Illegal " + cmd + " command: missing argument(s)");
}
inline int inumeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp) {
  char *end = nullptr;
  long v = strtol(str.c_str(), &end, 10);
  if (str.empty() || *end) lmp->error->all(file, line, "Expected integer parameter instead of '" + str + "' in input script or data file");
  return (int)v;
}
inline double numeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp) {
  char *end = nullptr;
  double v = strtod(str.c_str(), &end);
  if (str.empty() || *end) lmp->error->all(file, line, "Expected floating point parameter instead of '" + str + "' in input script or data file");
  return v;
}
}  // namespace utils

namespace MathConst {
static constexpr double MY_PI = 3.14159265358979323846;
static constexpr double MY_PIS = 1.77245385090551602729;
}

}  // namespace LAMMPS_NS

#define FixStyle(key, Class)   // style registration macro (style_fix.h machinery in upstream)

#endif
