// One subdomain's input data, the synthetic generator / partitioner, and the
// reference's on-disk formats.
//
//   SubMesh + read/write_input   the per-rank binary file of read_input_data /
//                                store_input_data_ (src/IO.cc:61-96, :99-127)
//   read/write_checking          the 2-line reference-norm file (src/IO.cc:26-39, :42-58)
//   generate_block               synthetic stand-in for the absent data/ tree
//                                (SURVEY.md §8d): structured Kuhn mesh, jittered,
//                                block-partitioned like the reference's pre-partitioned
//                                inputs (one subdomain per rank, 1-based interface lists)
#ifndef MFB_MESH_DATA_H
#define MFB_MESH_DATA_H

#include <cstdint>
#include <string>
#include <vector>

namespace mfb {

struct SubMesh {
    int nbElem = 0, nbNodes = 0, nbEdges = 0, nbIntf = 0, nbIntfNodes = 0, nbBoundNodes = 0;
    std::vector<double> coord;          // nbNodes * 3, AoS
    std::vector<int> elemToNode;        // nbElem * 4, 1-based local node ids
    std::vector<int> neighborsList;     // max(nbIntf,1) * 3, 1-based ranks (first nbIntf used)
    std::vector<int> intfIndex;         // nbIntf + 1
    std::vector<int> intfNodes;         // nbIntfNodes, 1-based local node ids
    std::vector<int> boundNodesCode;    // nbNodes
    std::vector<int64_t> globalNode;    // nbNodes, 0-based global id (generator only; not on disk)
};

// Structured grid of nx*ny*nz cubes, 6 tetrahedra per cube around the main diagonal,
// node positions jittered by +-0.1 cell (stateless hash of the global node id, so every
// rank agrees), boundary codes 52 (i==0), 53 (j==0), 54 (k==0), 10 (j==ny; wins).
// The px*py*pz block `rank` = (bz*py + by)*px + bx gets the cubes of its block.
// Returns 0, or -1 on invalid arguments.
int generate_block (int nx, int ny, int nz, int px, int py, int pz, int rank,
                    uint64_t seed, SubMesh &out);

// Largest px*py*pz <= maxRanks (each factor <= its axis) with the most cubic blocks.
void choose_blocks (int nx, int ny, int nz, int maxRanks, int &px, int &py, int &pz);

// <dataPath>/<mesh>/inputs/<op>_<nbBlocks>_<rank>
std::string input_path (const std::string &dataPath, const std::string &mesh,
                        const std::string &op, int nbBlocks, int rank);
// <dataPath>/<mesh>/checkings/<op>_<nbBlocks>_<rank>
std::string checking_path (const std::string &dataPath, const std::string &mesh,
                           const std::string &op, int nbBlocks, int rank);

int write_input (const std::string &file, const SubMesh &m);   // 0 / -1
int read_input (const std::string &file, SubMesh &m);          // 0 / -1
int write_checking (const std::string &file, double matrixNorm, double precNorm);
int read_checking (const std::string &file, double &matrixNorm, double &precNorm);

}  // namespace mfb

#endif
