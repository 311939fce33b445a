// TEST AID, not product code: compiles the SOURCE of the RING kernel (csrc/kernels_ring.cu) with g++
// and runs it on host threads — one std::thread per CUDA thread, CTAs one after the other — through
// tools/cuda_cta_emulation.h.  tests/test_ring_plan.py compares the result with the oracle; built
// with -fsanitize=thread (make -C mini-fem_b200 ringkernel-tsan) the same run checks the kernel's
// synchronisation protocol (buffer reuse across tiles, mbarrier phases, block barriers) for data
// races.  Built into tools/libmfb_ringkernel_host.so; nothing in libminifem_b200.so links it.
#define MFB_RING_HOST_EMULATION 1
#include "cuda_cta_emulation.h"

#include "../mini-fem_b200/csrc/kernels_ring.cu"

#include <string>

using namespace mfb;

static std::string g_error;

extern "C" const char *mfb_ring_kernel_host_error (void) { return g_error.c_str (); }

// One launch of ring_assembly_kernel over all tiles with `ctas` CTAs of 256 threads (the
// persistent-grid stride is `ctas`; threads = 384 selects the two-CTAs-per-SM instantiation).  fusePrec as
// on the device.  values / prec are host arrays.
extern "C" int mfb_ring_kernel_host (int operatorID, int nbNodes, int nbElem, const int *elemToNode, const int *row,
                                     const int *col, const double *coord, const int *checkBounds, const uint8_t *isInterface,
                                     int maxRows, int maxEntries, int ctas, int fusePrec, double *values, double *prec, int threads)
{
    RingPlanLimits lim;
    if (maxRows > 0) lim.maxRows = maxRows;
    if (maxEntries > 0) lim.maxEntries = maxEntries;
    RingPlan hp;
    if (build_ring_plan (nbNodes, nbElem, elemToNode, row, col, coord, isInterface, checkBounds, lim, hp, g_error) != 0) return -1;
    std::vector<uint64_t> packed (hp.tileOffset);
    for (int t = 0; t < hp.nbTiles; t++) packed[t] |= (uint64_t)(hp.header (t)->headBytes >> 4) << 48;
    // 16-byte aligned copy of the records, as cudaMalloc would give
    std::vector<uint8_t> blobStore (hp.blob.size () + 32);
    uint8_t *blob = reinterpret_cast<uint8_t*> (((uintptr_t)blobStore.data () + 15) & ~(uintptr_t)15);
    memcpy (blob, hp.blob.data (), hp.blob.size ());
    RingArgs args;
    args.plan.blob = blob; args.plan.tileOffset = packed.data ();
    args.plan.nbTiles = hp.nbTiles; args.plan.nbInterfaceTiles = hp.nbInterfaceTiles;
    args.plan.maxRows = std::max (hp.maxRows, 1); args.plan.maxNodes = std::max (hp.maxNodes, 4);
    args.plan.maxEntries = std::max (hp.maxEntries, 1);
    args.plan.maxHeadBytes = std::max (hp.maxHeadBytes, 16u); args.plan.maxTailBytes = std::max (hp.maxTailBytes, 16u);
    args.smem = ring_smem_layout (operatorID, args.plan);
    args.coord = coord; args.values = values; args.prec = prec;
    args.fusePrec = fusePrec; args.firstTile = 0; args.lastTile = hp.nbTiles;
    static unsigned intfDone;
    intfDone = 0;
    args.intfDone = isInterface ? &intfDone : nullptr;
    args.pollNs = 0;
    if (hp.nbTiles == 0) return 0;
    const size_t smem = ring_smem_bytes (operatorID, args.plan);
    auto launch = [&] (int firstTile, int nbTiles, int grid) {
        if (nbTiles <= 0) return;
        args.firstTile = firstTile; args.lastTile = firstTile + nbTiles;
        grid = std::max (1, std::min (grid, nbTiles));
        auto run = [&] (auto kernel, int t) { cta_emu::launch (grid, t, smem, [&] () { kernel (args); }); };
        const bool lap = operatorID == 0;
        switch (threads) {
        case 1024: if (lap) run (ring_assembly_kernel<1, 1024, 1>, 1024); else run (ring_assembly_kernel<9, 1024, 1>, 1024); break;
        case 896:  if (lap) run (ring_assembly_kernel<1, 896, 1>, 896);   else run (ring_assembly_kernel<9, 896, 1>, 896); break;
        case 768:  if (lap) run (ring_assembly_kernel<1, 768, 1>, 768);   else run (ring_assembly_kernel<9, 768, 1>, 768); break;
        case 640:  if (lap) run (ring_assembly_kernel<1, 640, 1>, 640);   else run (ring_assembly_kernel<9, 640, 1>, 640); break;
        default:   if (lap) run (ring_assembly_kernel<1, 384, 2>, 384);   else run (ring_assembly_kernel<9, 384, 2>, 384); break;
        }
    };
    const int grid = ctas > 0 ? ctas : 3;
    if (isInterface) {
        // as the fused multi-GPU iteration launches it (capi.cu do_iteration): the tiles that own interface
        // nodes first, then the interior tiles on a grid whose CTAs take four tiles each
        const int nIntf = hp.nbInterfaceTiles, nInterior = hp.nbTiles - nIntf;
        launch (0, nIntf, grid);
        launch (nIntf, nInterior, (nInterior + 3) / 4);
    }
    else launch (0, hp.nbTiles, grid);
    if (isInterface && (int)intfDone != hp.nbInterfaceTiles * ring_write_out_warps (operatorID, threads == 640 || threads == 768 || threads == 896 || threads == 1024 ? threads : 384)) {
        g_error = "interface signal: " + std::to_string (intfDone) + " arrivals for " + std::to_string (hp.nbInterfaceTiles) + " interface tiles";
        return -1;
    }
    return 0;
}
