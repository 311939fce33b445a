// TEST INFRASTRUCTURE — definitions for oracle/shim/DC.h (the DC-lib pieces the
// reference's REF / COLORING builds call).  See that header for the call sites each
// function serves and which semantics are [inferred].
#include <DC.h>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>
#if defined(__x86_64__)
#include <x86intrin.h>
#endif

static thread_local std::vector<uint64_t> tl_avg;
static thread_local int tl_next = 0;

static inline uint64_t cycles_now ()
{
#if defined(__x86_64__)
    return __rdtsc ();
#else
    return (uint64_t)std::chrono::steady_clock::now ().time_since_epoch ().count ();
#endif
}

static inline double seconds_now ()
{
    using namespace std::chrono;
    return duration<double> (steady_clock::now ().time_since_epoch ()).count ();
}

DC_timer::DC_timer () : t0_ (0), tSum_ (0), c0_ (0), cSum_ (0), tCount_ (0), cCount_ (0)
{
    id_ = tl_next++;
}
DC_timer::~DC_timer ()
{
    if ((int)tl_avg.size () <= id_) tl_avg.resize (id_ + 1, 0);
    tl_avg[id_] = get_avg_cycles ();
}
void DC_timer::start_time ()  { t0_ = seconds_now (); }
void DC_timer::stop_time ()   { tSum_ += seconds_now () - t0_; tCount_++; }
void DC_timer::reset_time ()  { tSum_ = 0; tCount_ = 0; }
double DC_timer::get_avg_time () { return tCount_ ? tSum_ / tCount_ : 0.0; }
void DC_timer::start_cycles () { c0_ = cycles_now (); }
void DC_timer::stop_cycles ()  { cSum_ += cycles_now () - c0_; cCount_++; }
void DC_timer::reset_cycles () { cSum_ = 0; cCount_ = 0; }
uint64_t DC_timer::get_avg_cycles () { return cCount_ ? cSum_ / cCount_ : 0; }

int minifem_ref_timer_count () { return (int)tl_avg.size (); }
uint64_t minifem_ref_timer_avg_cycles (int i) { return i < (int)tl_avg.size () ? tl_avg[i] : 0; }
void minifem_ref_forget_timers () { tl_avg.clear (); tl_next = 0; }

const char *minifem_ref_data_path ()
{
    const char *p = getenv ("MINIFEM_DATA_PATH");
    return p ? p : ".";
}

void DC_create_nodeToElem (index_t &nodeToElem, int *elemToNode, int nbElem,
                           int dimElem, int nbNodes)
{
    int *start = nodeToElem.index;
    memset (start, 0, sizeof (int) * (nbNodes + 1));
    for (int k = 0; k < nbElem * dimElem; k++) start[elemToNode[k]]++;   // 1-based id
    for (int n = 0; n < nbNodes; n++) start[n + 1] += start[n];
    std::vector<int> fill (start, start + nbNodes);
    for (int e = 0; e < nbElem; e++) {
        for (int k = 0; k < dimElem; k++) {
            int n = elemToNode[e * dimElem + k] - 1;
            nodeToElem.value[fill[n]++] = e;
        }
    }
}

void DC_create_elemToElem (list_t *elemToElem, index_t &nodeToElem, int *elemToNode,
                           int firstElem, int lastElem, int dimElem)
{
    std::vector<int> stamp (lastElem + 1, -1), tmp;
    for (int e = firstElem; e <= lastElem; e++) {
        tmp.clear ();
        for (int k = 0; k < dimElem; k++) {
            int n = elemToNode[e * dimElem + k] - 1;
            for (int p = nodeToElem.index[n]; p < nodeToElem.index[n + 1]; p++) {
                int other = nodeToElem.value[p];
                if (other == e || other < firstElem || other > lastElem) continue;
                if (stamp[other] != e) { stamp[other] = e; tmp.push_back (other); }
            }
        }
        elemToElem[e].size = (int)tmp.size ();
        elemToElem[e].list = new int [tmp.size () + 1];
        memcpy (elemToElem[e].list, tmp.data (), sizeof (int) * tmp.size ());
    }
}

void DC_create_permutation (int *perm, int *part, int size, int nbPart)
{
    std::vector<int> slot (nbPart + 1, 0);
    for (int i = 0; i < size; i++) slot[part[i] + 1]++;
    for (int p = 0; p < nbPart; p++) slot[p + 1] += slot[p];
    for (int i = 0; i < size; i++) perm[i] = slot[part[i]]++;
}

void DC_permute_int_2d_array (int *tab, int *perm, int nbItem, int dimItem, int offset)
{
    std::vector<int> old (tab, tab + (size_t)nbItem * dimItem);
    for (int i = 0; i < nbItem; i++) {
        for (int k = 0; k < dimItem; k++) {
            tab[(size_t)perm[i] * dimItem + k] = old[(size_t)i * dimItem + k] + offset;
        }
    }
}
