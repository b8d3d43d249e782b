// Host-side flattening of the pedigree into the index arrays the kernels consume.
//
// Replaces, once per (pedigree, genotyped id list, back):
//   PedigreeDAG.get_parents_depth / balance_nodes / get_subset per animal and per pair
//   (pedigree/pedigree_dag.py:124-168,221-231), BayesPedigreeNetworkModel.generateModel
//   (bayesnet/model.py:173-230) and correctMatrix's bookkeeping: partner_pairs, blankets,
//   per-generation gen_pairs, singlekids (correct_genotypes3.py:556-565,684-686,851-853).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/genofix_b200.h"

struct GfxBlanketSet {
  std::vector<int32_t> off;   // nb+1
  std::vector<int32_t> row;   // genotype row of the node or -1
  std::vector<int8_t> p1, p2, last;
  std::vector<uint64_t> kidmask;  // per node: bit mask of its children inside the blanket
  std::vector<int8_t> q0, q1; // focal local indices (q1 = -1 for a single focal animal)
  // ancestor slots of each focal animal in level order (slot 0/1 = sire/dam, slots 2s+2 / 2s+3 = parents of slot s):
  // genotype row, -1 = node without genotype, -2 = the other focal animal, -3 = no such node
  std::vector<int32_t> slotrow;   // count() * 2 * 14
  // bit s: the node at slot s must not be unobserved for plain peeling (several children in the blanket, or
  // parents beyond the slot tree); bit 15: plain peeling never applies to this focal animal
  std::vector<uint16_t> slotbad;  // count() * 2
  std::vector<uint8_t> tree;  // bit f: the ancestry of focal f is a plain tree (every ancestor has one child in the blanket)
  int max_nodes = 0;
  int max_width = 0;
  int count() const { return (int)off.size() - 1; }
};

struct GfxHostSchedule {
  int n_ped = 0, n_rows = 0, back = 2;
  // rank
  std::vector<int32_t> anc6;         // n_rows*6 rows of S,D,SS,SD,DS,DD or -1
  std::vector<uint8_t> rank_flags;   // bit0 has ancestors, bit1 needs the generic engine
  std::vector<int32_t> rank_blanket; // blanket id (depth `back`) or -1
  std::vector<int32_t> prow;         // n_rows*2 rows of the own parents (rank gate) or -1
  // genotyped children CSR per row, reference iteration order
  std::vector<int32_t> child_off, child_row, child_other;
  // pairs
  std::vector<int32_t> pair_sire_row, pair_dam_row, pair_blanket;
  std::vector<int32_t> gen_off;      // n_generations+1 offsets into job_pair
  std::vector<int32_t> job_pair;
  // per generation apply lists: focal rows with their (job-in-generation*2 + side) entries
  std::vector<int32_t> app_gen_off;  // n_generations+1 offsets into app_focal_row / app_focal_off
  std::vector<int32_t> app_focal_row;
  std::vector<int32_t> app_focal_off; // size app_focal_row.size()+1 (global offsets into app_entry)
  std::vector<int32_t> app_entry;
  std::vector<int32_t> app_order;     // per generation: focal indices (local to the generation), heaviest first
  // singles
  std::vector<int32_t> single_row, single_blanket, single_order;
  // trios
  std::vector<int32_t> trio_row;     // rows with both parents genotyped, grouped by (sire, dam) family
  std::vector<int32_t> fam_off;      // family f owns trio_row[fam_off[f] .. fam_off[f+1]); families ordered by sire
  GfxBlanketSet bl;
  gfx_schedule_info_t info{};
};

// Returns GFX_OK or an error code with `err` filled.
int gfx_build_schedule(const gfx_pedigree_desc* ped, GfxHostSchedule& out, std::string& err);
