// grid_index.cuh — K2: device-built uniform-grid spatial index that replaces
// pcl::KdTreeFLANN::setInputCloud (MO:1333-1334) for the radius-bounded 5-NN queries
// of cornerOptimization / surfOptimization (MO:1099-1101, MO:1181-1183).
//
// The reference only uses a 5-NN result when the 5th squared distance is < 1.0, so an
// exact search inside the ball of radius 1 m is equivalent to the exact kd-tree search.
// Cells are slightly larger than that radius (x1.001) so the 3x3x3 neighbourhood of the
// query's cell always contains the whole ball.  Cell ids are x-fastest, which makes the
// three x-neighbours of a (y,z) row ONE contiguous run in the re-ordered point array:
// a query reads 9 runs, not 27 cells.  Layout in HBM: `sorted` float4 {x,y,z,bits(original
// index)} in cell order (coalesced 16 B loads), `row_begin` int[rows+1] (exclusive scan of the
// points per (y,z) row) and `cell_begin` int[ncell+1] (exclusive scan of the per-cell counts),
// which is only materialised inside occupied rows: a 100k-point local map occupies ~1 % of its
// ~1M cells, so the build never touches (and the queries never read) the empty 99 %.
#pragma once
#include "common.cuh"

namespace llb {

struct GridDesc {            // device-resident
    int mn[3], mx[3];        // ordered-int encoded bounds (atomics), reset after use
    float org[3];
    float inv_cell;
    float cell;
    int dim[3];
    int ncell;
    int n;                   // points indexed
    unsigned ticket;         // last-CTA election inside the build kernels
};

struct MapIndexView {        // what the query kernels need (all device pointers)
    const float4 *sorted;
    const int *cell_begin;   // valid ONLY inside occupied rows (+ the entry right after such a row)
    const int *row_begin;    // [dimy*dimz + 1] exclusive scan of the points per (y,z) row: row r is empty iff
                             // row_begin[r] == row_begin[r+1]; a query looks here first and skips empty rows
    const GridDesc *desc;
};

struct GridJob {             // one index build: input cloud + the buffers of the GridIndex that receives it
    const float4 *pts; const int *n_dev; int n_host;
    GridDesc *desc; int *counts; int *cell_begin; int *cell_of; int *rank; int *row_cnt; int *row_begin; float4 *sorted;
};

class GridIndex {
public:
    // the tables are zeroed by kernels on `s`: pass the stream the builds will run on (a non-blocking stream does not
    // wait for work on the default stream)
    void init(int max_cells, cudaStream_t s = nullptr);
    void release();
    // Builds the indices of two maps with one set of five launches.  n_upper is a host upper
    // bound, n_dev (optional) the device-resident length.  Returns the number of launches.
    static int build_pair(GridIndex &a, const float4 *pa, const int *na_dev, int na_upper,
                          GridIndex &b, const float4 *pb, const int *nb_dev, int nb_upper, float radius, cudaStream_t s);
    // batched form: job() sizes this index's buffers for n_upper points and returns the job record; a
    // device-resident table of such records (any number of maps) is then built by ONE set of five launches
    GridJob job(const float4 *pts, const int *n_dev, int n_upper);
    static int build_table(const GridJob *table_dev, int count, int n_upper_max, float radius, int max_cells,
                           int ctas_per_map, cudaStream_t s);
    int max_cells() const { return max_cells_; }
    MapIndexView view() const { return MapIndexView{ sorted_.p, cell_begin_.p, row_begin_.p, desc_.p }; }
    const GridDesc *desc_dev() const { return desc_.p; }

private:
    int max_cells_ = 1 << 23;
    DevBuf<GridDesc> desc_;
    DevBuf<float4> sorted_;
    DevBuf<int> counts_;       // dense per-cell counts (kept zero between builds)
    DevBuf<int> cell_begin_;   // exclusive scan of the counts
    DevBuf<int> cell_of_;      // per point cell id
    DevBuf<int> rank_;         // per point rank inside its cell
    DevBuf<int> row_cnt_;      // points per (y,z) row (kept zero between builds)
    DevBuf<int> row_begin_;    // exclusive scan of the row counts
};

#ifdef __CUDACC__
__device__ __forceinline__ int grid_coord(float p, float org, float inv_cell)
{
    return (int)floorf((p - org) * inv_cell);
}
#endif

}  // namespace llb
