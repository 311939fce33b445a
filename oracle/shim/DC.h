// TEST INFRASTRUCTURE — stand-in for the un-vendored DC-lib header <DC.h>.
//
// The reference includes <DC.h> from main.cc:26, assembly.h:20, matrix.h:20,
// coloring.h:20, FEM.h:23 and halo.h:23, but DC-lib (EXAPARS/DC-lib, version
// unpinned: build/iMake:21-24 just points at $HOME/DC-lib) is not in the tree.
// This shim declares only what the REF / COLORING builds touch.  Semantics marked
// [inferred] are deduced from the reference's call sites and are this repository's
// definition of "bit-exact structure" ("parity unpinned").
//
// Only oracle/Makefile uses this file, to compile the reference's own sources from
// /root/reference into oracle/_ref/.  Nothing in the product includes it.
#ifndef MINIFEM_ORACLE_DC_SHIM_H
#define MINIFEM_ORACLE_DC_SHIM_H

#include <cstdint>
#include <string>

// main.cc:90 prints it; only the D&C builds give it meaning.
#define MAX_ELEM_PER_PART 0

// main.cc:118,243-244 / matrix.cc:63-68: CSR-like inverse map, 0-based offsets+ids.
typedef struct index_s { int *index, *value; } index_t;

// coloring.cc:57-58,93: neighbour list of one element.
typedef struct list_s { int *list; int size; list_s () : list (nullptr), size (0) {}
                        ~list_s () { delete[] list; } } list_t;

// Only named in signatures that the REF / COLORING builds compile out.
typedef struct DCargs_s DCargs_t;
typedef struct DCcommArgs_s DCcommArgs_t;

// FEM.cc:156,182-233 (cycles) and main.cc:117,143-151 (seconds).
class DC_timer {
public:
    DC_timer ();
    ~DC_timer ();
    void start_time ();
    void stop_time ();
    void reset_time ();
    double get_avg_time ();
    void start_cycles ();
    void stop_cycles ();
    void reset_cycles ();
    uint64_t get_avg_cycles ();
private:
    double   t0_, tSum_;
    uint64_t c0_, cSum_;
    int      tCount_, cCount_, id_;
};

// main.cc:247, coloring.cc:90.  [inferred] node -> incident elements, elements in
// increasing id, node slots 0-based (elemToNode holds 1-based ids).
void DC_create_nodeToElem (index_t &nodeToElem, int *elemToNode, int nbElem,
                           int dimElem, int nbNodes);

// coloring.cc:94.  [inferred] elements sharing >= 1 node with element i, for i in
// [firstElem, lastElem]; only the neighbour SET matters to the colouring.
void DC_create_elemToElem (list_t *elemToElem, index_t &nodeToElem, int *elemToNode,
                           int firstElem, int lastElem, int dimElem);

// coloring.cc:107.  [inferred] stable counting sort: perm[i] = destination slot.
void DC_create_permutation (int *perm, int *part, int size, int nbPart);

// main.cc:229.  [inferred] scatter rows: new[perm[i]] = old[i] + offset.
void DC_permute_int_2d_array (int *tab, int *perm, int nbItem, int dimItem, int offset);

// IO.cc:29,66 use the compile-time string macro DATA_PATH; the oracle build defines it
// as a call to this function so that the data tree can be chosen at run time.
const char *minifem_ref_data_path ();

// Average cycles of the timers that were created AND destroyed on the calling thread
// since the last minifem_ref_forget_timers(), indexed in construction order
// (FEM.cc:156 creates ASM, precInit, halo, precInver in that order).
int      minifem_ref_timer_count ();
uint64_t minifem_ref_timer_avg_cycles (int i);
void     minifem_ref_forget_timers ();

#endif
