// minifem_meshgen <dataPath> <LM6|EIB|FGN1|FGN4> <nx> <ny> <nz> <nbBlocks> [seed]
//
// Writes the synthetic stand-in for the absent data/ tree in the reference's own layout:
//   <dataPath>/<mesh>/inputs/{lap,ela}_<nbBlocks>_<rank>     (read_input_data, src/IO.cc:61-96)
// one file per rank of a px*py*pz = nbBlocks block partition (the most cubic one).  The
// matching <dataPath>/<mesh>/checkings/ files (src/IO.cc:26-39) are written by the driver:
//   MINIFEM_STORE_CHECKINGS=1 minifem_b200 <mesh> <op> <iterations>
// which plays the role of the reference's store_ref_assembly_ (src/IO.cc:42-58).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "minifem_b200.h"

int main (int argc, char **argv)
{
    if (argc < 7) {
        fprintf (stderr, "usage: %s <dataPath> <LM6|EIB|FGN1|FGN4> <nx> <ny> <nz> <nbBlocks> [seed]\n"
                         "  LM6-like: 25 25 40   EIB-like: 100 100 100\n", argv[0]);
        return EXIT_FAILURE;
    }
    const std::string dataPath = argv[1], mesh = argv[2];
    const int nx = atoi (argv[3]), ny = atoi (argv[4]), nz = atoi (argv[5]), nbBlocks = atoi (argv[6]);
    const unsigned long long seed = argc > 7 ? strtoull (argv[7], nullptr, 0) : 1;
    int px = 1, py = 1, pz = 1;
    mfb_choose_blocks (nx, ny, nz, nbBlocks, &px, &py, &pz);
    if (px * py * pz != nbBlocks) {
        fprintf (stderr, "Error: cannot cut %dx%dx%d cubes into %d blocks (closest: %dx%dx%d).\n", nx, ny, nz, nbBlocks, px, py, pz);
        return EXIT_FAILURE;
    }
    for (int rank = 0; rank < nbBlocks; rank++) {
        mfb_mesh *m = nullptr;
        if (mfb_mesh_generate (nx, ny, nz, px, py, pz, rank, seed, &m) != MFB_OK) {
            fprintf (stderr, "Error: %s\n", mfb_last_error ());
            return EXIT_FAILURE;
        }
        mfb_mesh_view v;
        mfb_mesh_get (m, &v);
        for (const char *op : {"lap", "ela"}) {
            const std::string file = dataPath + "/" + mesh + "/inputs/" + op + "_" + std::to_string (nbBlocks) + "_" + std::to_string (rank);
            if (mfb_mesh_write (m, file.c_str ()) != MFB_OK) {
                fprintf (stderr, "%s\n", mfb_last_error ());
                return EXIT_FAILURE;
            }
        }
        printf ("rank %d: %d elements, %d nodes, %d edges, %d interfaces (%d nodes)\n", rank, v.nbElem, v.nbNodes, v.nbEdges,
                v.nbIntf, v.nbIntfNodes);
        mfb_mesh_free (m);
    }
    printf ("wrote %s/%s/inputs/{lap,ela}_%d_<rank> (%dx%dx%d blocks)\n", dataPath.c_str (), mesh.c_str (), nbBlocks, px, py, pz);
    return EXIT_SUCCESS;
}
