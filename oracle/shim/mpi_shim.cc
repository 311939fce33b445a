// TEST INFRASTRUCTURE — thread-per-rank implementation of oracle/shim/mpi.h.
#include <mpi.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {

struct World {
    std::mutex mu;
    std::condition_variable cv;
    int nranks = 1;
    // (source, dest, tag) -> FIFO of buffered messages
    std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> box;
    // reduce rendezvous
    std::vector<std::vector<uint64_t>> contrib;
    int arrived = 0;
    long generation = 0;
} world;

thread_local int tl_rank = 0;

int type_bytes (MPI_Datatype t) { return (t == MPI_DOUBLE || t == MPI_UINT64_T) ? 8 : 1; }

}  // namespace

void minifem_ref_mpi_world (int nranks)
{
    std::lock_guard<std::mutex> lock (world.mu);
    world.nranks = nranks;
    world.box.clear ();
    world.contrib.assign (nranks, {});
    world.arrived = 0;
}

void minifem_ref_mpi_bind (int rank) { tl_rank = rank; }

int MPI_Init (int *, char ***) { return MPI_SUCCESS; }
int MPI_Finalize () { return MPI_SUCCESS; }
int MPI_Comm_size (MPI_Comm, int *size) { *size = world.nranks; return MPI_SUCCESS; }
int MPI_Comm_rank (MPI_Comm, int *rank) { *rank = tl_rank; return MPI_SUCCESS; }

int MPI_Irecv (void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm,
               MPI_Request *req)
{
    req->buf = buf; req->count = count; req->source = source; req->tag = tag;
    req->bytes = count * type_bytes (type);
    return MPI_SUCCESS;
}

int MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm)
{
    std::vector<char> msg ((size_t)count * type_bytes (type));
    memcpy (msg.data (), buf, msg.size ());
    {
        std::lock_guard<std::mutex> lock (world.mu);
        world.box[std::make_tuple (tl_rank, dest, tag)].push_back (std::move (msg));
    }
    world.cv.notify_all ();
    return MPI_SUCCESS;
}

int MPI_Waitall (int count, MPI_Request *reqs, MPI_Status *)
{
    for (int i = 0; i < count; i++) {
        std::unique_lock<std::mutex> lock (world.mu);
        auto key = std::make_tuple (reqs[i].source, tl_rank, reqs[i].tag);
        world.cv.wait (lock, [&] { return !world.box[key].empty (); });
        std::vector<char> &msg = world.box[key].front ();
        memcpy (reqs[i].buf, msg.data (),
                msg.size () < (size_t)reqs[i].bytes ? msg.size () : (size_t)reqs[i].bytes);
        world.box[key].pop_front ();
    }
    return MPI_SUCCESS;
}

int MPI_Reduce (const void *sendbuf, void *recvbuf, int count, MPI_Datatype, MPI_Op,
                int root, MPI_Comm)
{
    // Only FEM.cc:113 calls this: 4 x uint64, MPI_MAX, root 0.
    std::unique_lock<std::mutex> lock (world.mu);
    long gen = world.generation;
    const uint64_t *in = (const uint64_t*)sendbuf;
    world.contrib[tl_rank].assign (in, in + count);
    world.arrived++;
    if (world.arrived == world.nranks) {
        world.arrived = 0;
        world.generation++;
        world.cv.notify_all ();
    }
    else {
        world.cv.wait (lock, [&] { return world.generation != gen; });
    }
    if (tl_rank == root) {
        uint64_t *out = (uint64_t*)recvbuf;
        for (int k = 0; k < count; k++) {
            uint64_t m = 0;
            for (int r = 0; r < world.nranks; r++) {
                if (world.contrib[r][k] > m) m = world.contrib[r][k];
            }
            out[k] = m;
        }
    }
    return MPI_SUCCESS;
}
