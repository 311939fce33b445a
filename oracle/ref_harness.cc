// TEST INFRASTRUCTURE — C-ABI handle on the reference's OWN compiled sources.
//
// oracle/Makefile compiles /root/reference/src/{assembly,matrix,coloring,
// preconditioner,halo,FEM,IO,main}.cc where they lie (assembly.cc:120, Cilk Plus
// array notation, is rewritten on the fly by sed into the equivalent double loop)
// against oracle/shim/ and links this file in, giving
//     oracle/_ref/libminifem_ref_ref.so       (-DREF)
//     oracle/_ref/libminifem_ref_coloring.so  (-DCOLORING -DOMP, as shipped: the
//                                              per-colour omp pragma stays disabled)
// Every entry point below only forwards to a reference function; no arithmetic of the
// hot path is restated here.  Used by tests/ (to pin oracle/minifem_oracle.c and to
// generate tests/golden/) and by bench.py's reference arm.  Never by the product.
#include <mpi.h>
#include <DC.h>

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>

#include "globals.h"
#include "assembly.h"
#include "preconditioner.h"
#include "halo.h"
#include "matrix.h"
#include "coloring.h"
#include "FEM.h"

extern "C" {
void dqmrd4_ (int *nbNodes, int *boundNodesCode, int *nbBoundNodes,
              int *boundNodesList, int *error);
void e_essbcm_ (int *dimNode, int *nbNodes, int *nbBoundNodes, int *boundNodesList,
                int *boundNodesCode, int *checkBounds);
}
int minifem_ref_main (int argCount, char **argValue);   // main.cc built with -Dmain=...

namespace {
struct CoutSilencer {
    std::streambuf *saved;
    std::ostringstream sink;
    explicit CoutSilencer (bool on) : saved (nullptr) { if (on) saved = std::cout.rdbuf (sink.rdbuf ()); }
    ~CoutSilencer () { if (saved) std::cout.rdbuf (saved); }
};
}

extern "C" {

// 0 = REF build, 1 = COLORING build.
int mref_build_kind ()
{
#ifdef COLORING
    return 1;
#else
    return 0;
#endif
}

// main.cc:243-249: DC_create_nodeToElem + create_nodeToNode.  `col` must hold the
// nbEdges the input file announces (IO.cc:77); returns nodeToNodeRow[nbNodes].
int mref_create_nodeToNode (int *elemToNode, int nbElem, int nbNodes, int *row, int *col)
{
    index_t nodeToElem;
    nodeToElem.index = new int [nbNodes + 1];
    nodeToElem.value = new int [(size_t)nbElem * DIM_ELEM];
    DC_create_nodeToElem (nodeToElem, elemToNode, nbElem, DIM_ELEM, nbNodes);
    create_nodeToNode (row, col, nodeToElem, elemToNode, nbNodes);
    delete[] nodeToElem.value;
    delete[] nodeToElem.index;
    return row[nbNodes];
}

// Upper bound of nodeToNodeRow[nbNodes] so that callers can size `col` safely.
long mref_nodeToNode_capacity (int nbElem) { return (long)nbElem * DIM_ELEM * DIM_ELEM; }

// main.cc:325-327.
void mref_create_elemToEdge (int *row, int *col, int *elemToNode, int *elemToEdge,
                             int nbElem)
{
    create_elemToEdge (row, col, elemToNode, elemToEdge, nbElem);
}

// main.cc:216-230 (COLORING build): colour, then permute elemToNode in place.
// colorToElemOut needs 129 ints.  Returns nbTotalColors; the globals colorToElem /
// nbTotalColors stay set for assembly().  REF build: returns -1.
int mref_coloring (int *elemToNode, int nbElem, int nbNodes, int *colorPermOut,
                   int *colorToElemOut)
{
#ifdef COLORING
    delete[] colorToElem;
    colorToElem = nullptr;
    coloring_creation (elemToNode, colorPermOut, nbElem, nbNodes);
    DC_permute_int_2d_array (elemToNode, colorPermOut, nbElem, DIM_ELEM, 0);
    memcpy (colorToElemOut, colorToElem, sizeof (int) * (nbTotalColors + 1));
    return nbTotalColors;
#else
    (void)elemToNode; (void)nbElem; (void)nbNodes; (void)colorPermOut; (void)colorToElemOut;
    return -1;
#endif
}

// Installs colours computed elsewhere into the globals coloring_assembly reads (globals.h:43-44;
// assembly.cc:593-611), for callers that permuted elemToNode themselves.  REF build: no-op.
void mref_set_colors (const int *colorToElemIn, int nbColors)
{
#ifdef COLORING
    delete[] colorToElem;
    colorToElem = new int [nbColors + 1];
    memcpy (colorToElem, colorToElemIn, sizeof (int) * (nbColors + 1));
    nbTotalColors = nbColors;
#else
    (void)colorToElemIn; (void)nbColors;
#endif
}

// main.cc:340-346.
int mref_boundary_mask (int *boundNodesCode, int nbNodes, int nbBoundNodes,
                        int *checkBounds)
{
    int dimNode = DIM_NODE, error = 0, nb = nbBoundNodes;
    std::vector<int> list (nbBoundNodes > 0 ? nbBoundNodes : 1);
    dqmrd4_ (&nbNodes, boundNodesCode, &nb, list.data (), &error);
    e_essbcm_ (&dimNode, &nbNodes, &nb, list.data (), boundNodesCode, checkBounds);
    return error;
}

void mref_assembly (double *coord, double *values, int *row, int *col, int *elemToNode,
                    int *elemToEdge, int nbElem, int nbEdges, int operatorDim,
                    int operatorID)
{
    assembly (coord, values, row, col, elemToNode, elemToEdge, nbElem, nbEdges,
              operatorDim, operatorID);
}

void mref_prec_init (double *prec, double *values, int *row, int *col, int nbNodes,
                     int operatorDim)
{
    prec_init (prec, values, row, col, nbNodes, operatorDim);
}

void mref_prec_inversion (double *prec, int *row, int *col, int *checkBounds,
                          int nbNodes, int operatorID)
{
    prec_inversion (prec, row, col, checkBounds, nbNodes, operatorID);
}

double mref_norm (double *tab, int size) { return compute_double_norm (tab, size); }

typedef struct {
    double *coord, *values, *prec;
    int *row, *col, *elemToNode, *elemToEdge, *intfIndex, *intfNodes, *neighborsList,
        *checkBounds;
    int nbElem, nbNodes, nbEdges, nbIntf, nbIntfNodes;
} mref_rank_t;

// FEM.cc:139-285, one thread per subdomain.  cyclesOut[4] = per-stage averages over
// iterations 1..n-1 (FEM.cc:182), max over ranks (FEM.cc:113).  *tscHz = measured
// TSC rate so that callers can turn cycles into seconds.  Returns 0.
int mref_fem_loop (int nranks, mref_rank_t *ranks, int nbIter, int operatorID,
                   uint64_t *cyclesOut, double *tscHz, int verbose)
{
    int operatorDim = operatorID == 0 ? 1 : DIM_NODE * DIM_NODE;
    minifem_ref_mpi_world (nranks);
    std::vector<uint64_t> perRank ((size_t)nranks * 4, 0);
    CoutSilencer quiet (!verbose);

    DC_timer wall;
    auto t0 = std::chrono::steady_clock::now ();
    wall.start_cycles ();
    auto body = [&] (int r) {
        minifem_ref_mpi_bind (r);
        minifem_ref_forget_timers ();
        mref_rank_t &d = ranks[r];
        FEM_loop (d.prec, d.coord, d.values, d.row, d.col, d.elemToNode, d.elemToEdge,
                  d.intfIndex, d.intfNodes, d.neighborsList, d.checkBounds, d.nbElem,
                  d.nbNodes, d.nbEdges, d.nbIntf, d.nbIntfNodes, nbIter, nranks, r,
                  operatorDim, operatorID);
        for (int k = 0; k < 4; k++) perRank[(size_t)r * 4 + k] = minifem_ref_timer_avg_cycles (k);
    };
    if (nranks == 1) {
        body (0);
    }
    else {
        std::vector<std::thread> pool;
        for (int r = 0; r < nranks; r++) pool.emplace_back (body, r);
        for (auto &t : pool) t.join ();
    }
    wall.stop_cycles ();
    double secs = std::chrono::duration<double> (std::chrono::steady_clock::now () - t0).count ();
    *tscHz = secs > 0 ? (double)wall.get_avg_cycles () / secs : 0.0;
    for (int k = 0; k < 4; k++) {
        uint64_t m = 0;
        for (int r = 0; r < nranks; r++) if (perRank[(size_t)r * 4 + k] > m) m = perRank[(size_t)r * 4 + k];
        cyclesOut[k] = m;
    }
    return 0;
}

// The reference's whole driver (main.cc:96-388), single rank: reads
// $MINIFEM_DATA_PATH/<mesh>/inputs/<op>_1_0 and .../checkings/<op>_1_0, writes
// numerical_results_0 in the current directory.
int mref_main (const char *mesh, const char *op, const char *nbIter)
{
    minifem_ref_mpi_world (1);
    minifem_ref_mpi_bind (0);
    char a0[] = "miniFEM", *argv[5];
    std::string m (mesh), o (op), n (nbIter);
    argv[0] = a0; argv[1] = &m[0]; argv[2] = &o[0]; argv[3] = &n[0]; argv[4] = nullptr;
    return minifem_ref_main (4, argv);
}

}  // extern "C"
