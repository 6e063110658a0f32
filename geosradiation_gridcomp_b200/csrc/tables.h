// Host-side table initialisation of the RRTMG LW/SW path: reads the extracted reference data
// blob (tools/extract_tables.py) and performs what rrtmg_lw_ini / rrtmg_sw_ini do once per
// Initialize (LW/src/rrtmg_lw_init.F90:22-165, SW/src/rrtmg_sw_init.F90:49-171): the
// 256->140 and 224->112 g-point reductions, the tau/exp/tfn lookup tables and the band maps.
// Reduced tables are stored g-point FASTEST ([lead][ng]) because the kernels loop over the
// g-points of a band inside one thread and read consecutive g from one table row.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace rrtmgx {

struct TableRef { size_t off = 0; size_t n = 0; bool ok() const { return n != 0; } };

struct HostTables {
    std::vector<double> arena;                 // every fp64 table, uploaded as one buffer
    std::map<std::string, TableRef> index;     // "lw.03.absa", "sw.17.sfluxref", "lw.exp_tbl", ...
    int lw_ngc[16], lw_ngs[16], lw_ngb[140], lw_nspa[16], lw_nspb[16];
    int sw_ngc[14], sw_ngs[14], sw_ngb[112], sw_nspa[14], sw_nspb[14], sw_icxa[14];
    int sw_nfor[14], sw_nsrc[14];
    double lw_delwave[16];
    double sw_rayl_scalar[14];
    bool sw_has_raylv[14];

    int load(const std::string &blob_path);    // 0 or RRTMGX_EBLOB
    TableRef find(const std::string &name) const;
    const double *ptr(const TableRef &r) const { return r.ok() ? arena.data() + r.off : nullptr; }
    TableRef add(const std::string &name, const std::vector<double> &v);
};

}  // namespace rrtmgx
