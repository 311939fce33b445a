// minifem_b200 $USE_CASE $OPERATOR $NB_ITERATIONS — the Mini-FEM driver on the B200 path.
//
// Host-side mirror of src/main.cc:96-388 (argument check, setup sequence, allocation),
// src/FEM.cc:139-285 (FEM_loop: four stages per iteration, cycle timers that skip
// iteration 0, max over ranks, "Average cycles" table) and src/FEM.cc:59-98
// (check_results against the per-mesh "checkings" file).  The stages themselves run on
// the GPU through the C ABI (include/minifem_b200.h); nothing here computes on the CPU.
//
// Compile-time switches of the reference become environment variables:
//   MINIFEM_DATA_PATH   DATA_PATH of build/iMake:20            (default ./data)
//   MINIFEM_PATH        ring | tiled | atomic | color          (default ring; color = the
//                       COLORING build: colour + permute before the CSR, main.cc:209-236)
//   MINIFEM_FUSED       1 = one fused launch per iteration, reported like the
//                       multithreaded-comm build reports (FEM.cc:179-180: everything under
//                       "Matrix assembly")                     (default 0: four timed stages)
//   MINIFEM_DEVICE      CUDA device ordinal                    (default LOCAL_RANK or 0)
//   MINIFEM_GPU_SETUP   1 = colouring and CSR built on the device (same layouts, bit for bit)
//   MINIFEM_DEVICE_NORMS 1 = the two check norms are reduced on the device (default: host, serial sum)
//   MINIFEM_STORE_CHECKINGS  1 = write the checkings file from this run's norms instead of
//                       comparing with it (the role of store_ref_assembly_, src/IO.cc:42-58)
//   RANK / WORLD_SIZE   MPI rank / size of the reference (one process per GPU);
//   MINIFEM_RENDEZVOUS  directory shared by the ranks (NCCL id + timer reduction)
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <thread>
#include <sys/stat.h>
#include <ctime>
#include <vector>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif

#include "minifem_b200.h"

using namespace std;

namespace {

string meshName, operatorName;            // globals.h:41

uint64_t cycles_now ()
{
#if defined(__x86_64__)
    return __rdtsc ();
#else
    return (uint64_t)chrono::steady_clock::now ().time_since_epoch ().count ();
#endif
}

// DC-lib's DC_timer as FEM.cc / main.cc use it: averages of seconds or cycles.
struct Timer {
    double t0 = 0, tSum = 0;
    uint64_t c0 = 0, cSum = 0;
    int tCount = 0, cCount = 0;
    static double now () { return chrono::duration<double> (chrono::steady_clock::now ().time_since_epoch ()).count (); }
    void start_time () { t0 = now (); }
    void stop_time () { tSum += now () - t0; tCount++; }
    void reset_time () { tSum = 0; tCount = 0; }
    double get_avg_time () const { return tCount ? tSum / tCount : 0; }
    void start_cycles () { c0 = cycles_now (); }
    void stop_cycles () { cSum += cycles_now () - c0; cCount++; }
    uint64_t get_avg_cycles () const { return cCount ? cSum / cCount : 0; }
};

[[noreturn]] void die (const string &msg)
{
    cerr << msg << "\n";
    exit (EXIT_FAILURE);
}

void check (int rc, const char *what)
{
    if (rc != MFB_OK) die (string ("Error: ") + what + ": " + mfb_last_error ());
}

void help ()
{
    cerr << "Please specify:\n"
         << " 1. The test case: LM6, EIB or FGN.\n"
         << " 2. The operator: lap or ela.\n"
         << " 3. The number of iterations.\n";
}

// main.cc:57-94
void check_args (int argCount, char **argValue, int *nbIter, int rank)
{
    if (argCount < 4) {
        if (rank == 0) help ();
        exit (EXIT_FAILURE);
    }
    meshName = argValue[1];
    if (meshName != "LM6" && meshName != "EIB" && meshName != "FGN1" && meshName != "FGN4") {
        if (rank == 0) { cerr << "Incorrect argument \"" << meshName << "\".\n"; help (); }
        exit (EXIT_FAILURE);
    }
    operatorName = argValue[2];
    if (operatorName != "lap" && operatorName != "ela") {
        if (rank == 0) { cerr << "Incorrect argument \"" << operatorName << "\".\n"; help (); }
        exit (EXIT_FAILURE);
    }
    *nbIter = (int)strtol (argValue[3], nullptr, 0);
    if (*nbIter < 1) {
        if (rank == 0) cerr << "Number of iterations must be at least 1.\n";
        exit (EXIT_FAILURE);
    }
    if (rank == 0) {
        cout << "\t\t* Mini-FEM *\n\n"
             << "Test case              : \"" << meshName << "\"\n"
             << "Operator               : \"" << operatorName << "\"\n"
             << "Elements per partition :  " << 0 << "\n"
             << "Iterations             :  " << *nbIter << "\n\n"
             << scientific << setprecision (1);
    }
}

int env_int (const char *name, int fallback)
{
    const char *v = getenv (name);
    return v ? atoi (v) : fallback;
}

string env_str (const char *name, const string &fallback)
{
    const char *v = getenv (name);
    return v ? string (v) : fallback;
}

// Rendezvous files of one run: named after a token the ranks of a run share (MINIFEM_RUN_TOKEN, else torchrun's
// run id and port), never older than this process (a file left behind by a crashed run is ignored), and
// removed by rank 0 when the run ends.
const time_t kProcessStart = time (nullptr);

string run_token ()
{
    const char *t = getenv ("MINIFEM_RUN_TOKEN");
    if (t) return string ("_") + t;
    string token;
    if (const char *id = getenv ("TORCHELASTIC_RUN_ID")) token += string ("_") + id;
    if (const char *port = getenv ("MASTER_PORT")) token += string ("_") + port;
    return token;
}

void wait_for_file (const string &path)
{
    for (int tries = 0; tries < 60000; tries++) {
        struct stat st;
        if (stat (path.c_str (), &st) == 0 && st.st_mtime + 2 >= kProcessStart) {
            ifstream f (path, ios::binary);
            if (f.good ()) return;
        }
        this_thread::sleep_for (chrono::milliseconds (5));
    }
    die ("Error: timed out waiting for " + path);
}

// FEM.cc:101-136 (MPI_Reduce MAX to rank 0 becomes a file rendezvous)
void get_average_cycles (const Timer &asmT, const Timer &initT, const Timer &haloT, const Timer &invT,
                         int nbBlocks, int rank, const string &rendezvous)
{
    uint64_t local[4] = {asmT.get_avg_cycles (), initT.get_avg_cycles (), haloT.get_avg_cycles (),
                         invT.get_avg_cycles ()}, global[4];
    memcpy (global, local, sizeof local);
    if (nbBlocks > 1) {
        const string mine = rendezvous + "/cycles" + run_token () + "_" + to_string (rank);
        { ofstream f (mine + ".tmp", ios::binary); f.write ((const char*)local, sizeof local); }
        rename ((mine + ".tmp").c_str (), mine.c_str ());
        if (rank == 0) {
            for (int r = 1; r < nbBlocks; r++) {
                const string theirs = rendezvous + "/cycles" + run_token () + "_" + to_string (r);
                wait_for_file (theirs);
                uint64_t other[4];
                { ifstream f (theirs, ios::binary); f.read ((char*)other, sizeof other); }
                for (int k = 0; k < 4; k++) if (other[k] > global[k]) global[k] = other[k];
                remove (theirs.c_str ());
            }
            remove (mine.c_str ());
            remove ((rendezvous + "/nccl_id" + run_token ()).c_str ());      // every rank has initialised its communicator long ago
            for (int r = 0; r < nbBlocks; r++) {                             // ... and mapped its neighbours' windows
                remove ((rendezvous + "/p2p_card" + run_token () + "_" + to_string (r)).c_str ());
                remove ((rendezvous + "/p2p_ok" + run_token () + "_" + to_string (r)).c_str ());
            }
        }
    }
    if (rank == 0) {
        cout << "Average cycles\n";
        cout << "----------------------------------------------\n";
        cout << "  Matrix assembly               : " << global[0] << endl;
        cout << "  Preconditioner initialization : " << global[1] << endl;
        cout << "  Halo exchange                 : " << global[2] << endl;
        cout << "  Preconditioner inversion      : " << global[3] << endl;
        cout << "  Total                         : " << global[0] + global[1] + global[2] + global[3] << endl;
        cout << "----------------------------------------------\n\n";
    }
}

// FEM.cc:139-285
void FEM_loop (mfb_ctx *ctx, int nbIter, int nbBlocks, int rank, bool fused, const string &rendezvous)
{
    Timer ASMtimer, precInitTimer, haloTimer, precInverTimer;
    for (int iter = 0; iter < nbIter; iter++) {
        const bool timed = nbIter == 1 || iter > 0;
        if (rank == 0) cout << iter << ". Matrix assembly...                ";
        if (timed) ASMtimer.start_cycles ();
        check (fused ? mfb_ctx_iteration (ctx) : mfb_ctx_assembly (ctx), "assembly");
        check (mfb_ctx_sync (ctx), "assembly");
        if (timed) ASMtimer.stop_cycles ();
        if (rank == 0) cout << "done\n";

        if (rank == 0) cout << "   Preconditioner initialization...  ";
        if (timed) precInitTimer.start_cycles ();
        if (!fused) { check (mfb_ctx_prec_init (ctx), "prec_init"); check (mfb_ctx_sync (ctx), "prec_init"); }
        if (timed) precInitTimer.stop_cycles ();
        if (rank == 0) cout << "done\n";

        if (rank == 0) cout << "   Halo exchange...                  ";
        if (timed) haloTimer.start_cycles ();
        if (!fused) { check (mfb_ctx_halo_exchange (ctx), "halo exchange"); check (mfb_ctx_sync (ctx), "halo exchange"); }
        if (timed) haloTimer.stop_cycles ();
        if (rank == 0) cout << "done\n";

        if (rank == 0) cout << "   Preconditioner inversion...       ";
        if (timed) precInverTimer.start_cycles ();
        if (!fused) { check (mfb_ctx_prec_inversion (ctx), "prec_inversion"); check (mfb_ctx_sync (ctx), "prec_inversion"); }
        if (timed) precInverTimer.stop_cycles ();
        if (rank == 0) cout << "done\n\n";
    }
    get_average_cycles (ASMtimer, precInitTimer, haloTimer, precInverTimer, nbBlocks, rank, rendezvous);

    float ms[5];
    check (mfb_ctx_stage_ms (ctx, ms), "stage timings");
    if (rank == 0) {
        cout << "GPU time of the last iteration (ms, CUDA events)\n"
             << "----------------------------------------------\n" << fixed << setprecision (4);
        if (fused) cout << "  Fused iteration               : " << ms[4] << "\n";
        else cout << "  Matrix assembly               : " << ms[0] << "\n"
                  << "  Preconditioner initialization : " << ms[1] << "\n"
                  << "  Halo exchange                 : " << ms[2] << "\n"
                  << "  Preconditioner inversion      : " << ms[3] << "\n";
        cout << "----------------------------------------------\n\n" << scientific << setprecision (1);
    }
}

// FEM.cc:59-98
void check_results (double matrixNorm, double precNorm, int nbBlocks, int rank, const string &dataPath)
{
    double refMatrixNorm, refPrecNorm;
    const string file = dataPath + "/" + meshName + "/checkings/" + operatorName + "_" + to_string (nbBlocks) +
                        "_" + to_string (rank);
    if (env_int ("MINIFEM_STORE_CHECKINGS", 0)) {              // store_ref_assembly_, IO.cc:42-58
        if (mfb_checking_write (file.c_str (), matrixNorm, precNorm) != MFB_OK) die (mfb_last_error ());
        if (rank == 0) cout << "Stored reference checking: " << file << endl << endl;
        return;
    }
    if (mfb_checking_read (file.c_str (), &refMatrixNorm, &refPrecNorm) != MFB_OK) die (mfb_last_error ());
    auto report = [&] (ostream &o, int r) {
        o << "Numerical stability of rank " << r << endl
          << "----------------------------------------------" << endl
          << "  Matrix -> reference norm : " << refMatrixNorm << endl
          << "              current norm : " << matrixNorm << endl
          << "                difference : " << abs (refMatrixNorm - matrixNorm) / refMatrixNorm << endl << endl
          << "    Prec -> reference norm : " << refPrecNorm << endl
          << "              current norm : " << precNorm << endl
          << "                difference : " << abs (refPrecNorm - precNorm) / refPrecNorm << endl
          << "----------------------------------------------" << endl;
    };
    ofstream resultFile ("numerical_results_" + to_string (rank), ios::out | ios::trunc);
    report (resultFile, rank);
    resultFile.close ();
    if (rank == 0) {
        report (cout, 0);
        cout << "(see numerical_results files for all ranks)" << endl << endl;
    }
}

}  // namespace

int main (int argCount, char **argValue)
{
    // Process initialization (main.cc:98-114): one process per GPU
    const int nbBlocks = max (env_int ("WORLD_SIZE", 1), 1), rank = env_int ("RANK", 0);
    const string dataPath = env_str ("MINIFEM_DATA_PATH", "./data");
    const string pathName = env_str ("MINIFEM_PATH", "ring");
    const string rendezvous = env_str ("MINIFEM_RENDEZVOUS", ".");
    const bool fused = env_int ("MINIFEM_FUSED", 0) != 0;
    const int device = env_int ("MINIFEM_DEVICE", env_int ("LOCAL_RANK", 0));
    const bool gpuSetup = env_int ("MINIFEM_GPU_SETUP", 0) != 0;
    int path = MFB_PATH_RING;
    if (pathName == "atomic") path = MFB_PATH_ATOMIC;
    else if (pathName == "color") path = MFB_PATH_COLOR;
    else if (pathName == "tiled") path = MFB_PATH_TILED;
    else if (pathName != "ring") die ("Incorrect MINIFEM_PATH \"" + pathName + "\" (tiled, atomic, color or ring).");

    Timer timer;
    int nbIter;
    check_args (argCount, argValue, &nbIter, rank);
    const int operatorID = operatorName == "lap" ? 0 : 1;          // main.cc:131-138
    const int operatorDim = operatorID == 0 ? 1 : 9;

    auto begin_step = [&] (const char *label) { if (rank == 0) { cout << label; cout.flush (); timer.start_time (); } };
    auto end_step = [&] () {
        if (rank == 0) {
            timer.stop_time ();
            cout << "done  (" << timer.get_avg_time () << " seconds)\n";
            timer.reset_time ();
        }
    };

    // Get the input data (main.cc:140-152)
    begin_step ("Reading input data...                ");
    const string inputFile = dataPath + "/" + meshName + "/inputs/" + operatorName + "_" + to_string (nbBlocks) +
                             "_" + to_string (rank);
    mfb_mesh *mesh = nullptr;
    if (mfb_mesh_read (inputFile.c_str (), &mesh) != MFB_OK) die (mfb_last_error ());
    mfb_mesh_view in;
    check (mfb_mesh_get (mesh, &in), "mesh view");
    end_step ();

    // Mesh coloring version (main.cc:209-236)
    vector<int> colorToElem (129, 0);
    int nbTotalColors = 0;
    if (path == MFB_PATH_COLOR) {
        begin_step ("Coloring of the mesh...              ");
        vector<int> colorPerm (max (in.nbElem, 1)), colorPart (max (in.nbElem, 1));
        const int rcColor = gpuSetup
            ? mfb_device_coloring_creation (in.elemToNode, in.nbElem, in.nbNodes, colorPart.data (), colorToElem.data (),
                                            colorPerm.data (), &nbTotalColors, device)
            : mfb_coloring_creation (in.elemToNode, in.nbElem, in.nbNodes, colorPart.data (), colorToElem.data (),
                                     colorPerm.data (), &nbTotalColors);
        if (rcColor != MFB_OK) die (mfb_last_error ());
        end_step ();
        begin_step ("Applying permutation...              ");
        check (mfb_permute_int_2d (in.elemToNode, colorPerm.data (), in.nbElem, 4), "permutation");
        end_step ();
    }

    // Create the CSR matrix (main.cc:238-255); nbEdges comes from the file header (IO.cc:77)
    begin_step ("Creating CSR matrix...               ");
    vector<int> nodeToNodeRow ((size_t)in.nbNodes + 1), nodeToNodeColumn (max (in.nbEdges, 1));
    int nbEdges = 0;
    if (gpuSetup) {                                            // MINIFEM_GPU_SETUP=1: the same layouts, built on the device
        const int rcCsr = mfb_device_create_nodeToNode (in.elemToNode, in.nbElem, in.nbNodes, nodeToNodeRow.data (),
                                                        nodeToNodeColumn.data (), in.nbEdges, &nbEdges, device);
        if (nbEdges != in.nbEdges) {
            die ("Error: the input file announces " + to_string (in.nbEdges) + " edges, the mesh has " + to_string (nbEdges) + ".");
        }
        check (rcCsr, "create_nodeToNode");
    }
    else {
        const int64_t counted = mfb_count_edges (in.elemToNode, in.nbElem, in.nbNodes);
        if (counted != in.nbEdges) {
            die ("Error: the input file announces " + to_string (in.nbEdges) + " edges, the mesh has " + to_string (counted) + ".");
        }
        check (mfb_create_nodeToNode (in.elemToNode, in.nbElem, in.nbNodes, nodeToNodeRow.data (),
                                      nodeToNodeColumn.data (), &nbEdges), "create_nodeToNode");
    }
    end_step ();

    // Compute the boundary conditions (main.cc:335-351)
    begin_step ("Computing boundary conditions...     ");
    vector<int> checkBounds ((size_t)max (in.nbNodes, 1) * 3);
    check (mfb_boundary_mask (in.boundNodesCode, in.nbNodes, checkBounds.data (), nullptr), "boundary mask");
    end_step ();

    // Device context: uploads the arrays, builds elemToEdge (main.cc:319-333) or the tile plan
    begin_step ("Preparing the GPU...                 ");
    mfb_problem prob;
    memset (&prob, 0, sizeof prob);
    prob.operatorID = operatorID;
    prob.nbElem = in.nbElem; prob.nbNodes = in.nbNodes; prob.nbEdges = nbEdges;
    prob.coord = in.coord; prob.elemToNode = in.elemToNode;
    prob.nodeToNodeRow = nodeToNodeRow.data (); prob.nodeToNodeColumn = nodeToNodeColumn.data ();
    prob.checkBounds = checkBounds.data ();
    prob.colorToElem = path == MFB_PATH_COLOR ? colorToElem.data () : nullptr;
    prob.nbTotalColors = nbTotalColors;
    prob.nbBlocks = nbBlocks; prob.rank = rank;
    prob.nbIntf = in.nbIntf; prob.nbIntfNodes = in.nbIntfNodes;
    prob.intfIndex = in.intfIndex; prob.intfNodes = in.intfNodes; prob.neighborsList = in.neighborsList;
    mfb_options opt;
    memset (&opt, 0, sizeof opt);
    opt.path = path; opt.device = device; opt.useGraph = path == MFB_PATH_COLOR;
    mfb_ctx *ctx = nullptr;
    check (mfb_ctx_create (&prob, &opt, &ctx), "GPU context");
    if (nbBlocks > 1) {
        unsigned char id[MFB_COMM_ID_BYTES];
        // the file names carry a per-run token (torchrun's run id / rendezvous port) so that a second run in
        // the same directory never reads the files of the first; rank 0 removes them when it is done
        const string idFile = rendezvous + "/nccl_id" + run_token ();
        if (rank == 0) {
            remove (idFile.c_str ());                                  // a leftover of an interrupted run
            check (mfb_comm_unique_id (id), "NCCL id");
            { ofstream f (idFile + ".tmp", ios::binary); f.write ((const char*)id, sizeof id); }
            rename ((idFile + ".tmp").c_str (), idFile.c_str ());
        }
        else {
            wait_for_file (idFile);
            ifstream f (idFile, ios::binary);
            f.read ((char*)id, sizeof id);
        }
        check (mfb_ctx_comm_init (ctx, id), "NCCL communicator");
        // Peer-to-peer windows for the fused RING iteration (mfb_ctx_p2p_*; the GASPI segments of main.cc under
        // -DGASPI): cards and the outcome travel through the rendezvous directory; every rank switches or none does.
        if (path == MFB_PATH_RING && fused && env_str ("MFB_HALO", "p2p") != "nccl") {
            auto put = [&] (const string &name, const void *data, size_t bytes) {
                { ofstream f (name + ".tmp", ios::binary); f.write ((const char*)data, (streamsize)bytes); }
                rename ((name + ".tmp").c_str (), name.c_str ());
            };
            const string cardBase = rendezvous + "/p2p_card" + run_token () + "_", okBase = rendezvous + "/p2p_ok" + run_token () + "_";
            vector<unsigned char> cards ((size_t)nbBlocks * MFB_P2P_CARD_BYTES, 0);
            char ok = mfb_ctx_p2p_card (ctx, cards.data () + (size_t)rank * MFB_P2P_CARD_BYTES) == MFB_OK;
            put (cardBase + to_string (rank), cards.data () + (size_t)rank * MFB_P2P_CARD_BYTES, MFB_P2P_CARD_BYTES);
            for (int r = 0; r < nbBlocks; r++) {
                if (r == rank) continue;
                wait_for_file (cardBase + to_string (r));
                ifstream f (cardBase + to_string (r), ios::binary);
                f.read ((char*)cards.data () + (size_t)r * MFB_P2P_CARD_BYTES, MFB_P2P_CARD_BYTES);
            }
            if (ok) ok = mfb_ctx_p2p_connect (ctx, cards.data ()) == MFB_OK;
            put (okBase + to_string (rank), &ok, 1);
            bool all = ok != 0;
            for (int r = 0; r < nbBlocks; r++) {
                if (r == rank) continue;
                wait_for_file (okBase + to_string (r));
                ifstream f (okBase + to_string (r), ios::binary);
                char theirs = 0;
                f.read (&theirs, 1);
                all = all && theirs != 0;
            }
            if (!all && ok) check (mfb_ctx_p2p_enable (ctx, 0), "peer-to-peer exchange off");
            if (rank == 0) cout << "Interface sum: " << (all ? "peer-to-peer windows (NVLink stores + epoch flags)" : "NCCL send / recv") << "\n";
        }
    }
    end_step ();

    // Main loop with assembly, solver & update (main.cc:353-367)
    if (rank == 0) cout << "\nMain FEM loop\n";
    FEM_loop (ctx, nbIter, nbBlocks, rank, fused, rendezvous);
    // The two norms check_results compares (FEM.cc:68-76).  Default: the arrays come back to the
    // host and are summed serially like compute_double_norm, so that the figures agree with the
    // reference's to the last digit; MINIFEM_DEVICE_NORMS=1 reduces them on the device instead.
    double matrixNorm = 0, precNorm = 0;
    if (env_int ("MINIFEM_DEVICE_NORMS", 0)) {
        check (mfb_ctx_norms (ctx, &matrixNorm, &precNorm), "norms");
    }
    else {
        vector<double> nodeToNodeValue ((size_t)max (nbEdges, 1) * operatorDim), prec ((size_t)max (in.nbNodes, 1) * operatorDim);
        check (mfb_ctx_download (ctx, nodeToNodeValue.data (), prec.data ()), "download");
        matrixNorm = mfb_double_norm (nodeToNodeValue.data (), (int64_t)nbEdges * operatorDim);
        precNorm = mfb_double_norm (prec.data (), (int64_t)in.nbNodes * operatorDim);
    }
    mfb_ctx_destroy (ctx);
    mfb_mesh_free (mesh);

    // Check matrix & prec arrays (main.cc:375-378)
    check_results (matrixNorm, precNorm, nbBlocks, rank, dataPath);
    return EXIT_SUCCESS;
}
