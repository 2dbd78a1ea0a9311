// vector_index.hpp -- C++ host-side mirror of NesterovskyBros.VectorIndex over the libvi_b200 C ABI (header only).
// Same names and error behaviour as the reference's C# surface for the split-tree path:
//   IndexBuilder::Build(points, storeFactory)  VectorIndex/IndexBuilder.cs:23-25
//   IRangeStore / MemoryRangeStore             VectorIndex/IRangeStore.cs:6-22, MemoryRangeStore.cs:7-32
//   RangeValue {Dimension, Mid, Id}            VectorIndex/RangeValue.cs:6-22
//   VectorIndex::Find(vector, distance, pred)  shape of MemoryVectorIndex.cs:242-245 over dbo.Search (DDL.sql:234-295)
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "vi_b200.h"

namespace NesterovskyBros::VectorIndex
{
struct RangeValue
{
  int32_t Dimension = 0;  // -1 = leaf; VI_DIM_NULL (-3) = null (VI_MODE_SQL, DDL.sql:193)
  float Mid = 0.0f;
  int64_t Id = 0;
};

using Point = std::pair<int64_t, std::vector<float>>;

struct IRangeStore
{
  virtual ~IRangeStore() = default;
  virtual void Add(int64_t id, const std::vector<float>& vector) = 0;
  virtual const std::vector<Point>& GetPoints() const = 0;
};

class MemoryRangeStore : public IRangeStore
{
 public:
  void Add(int64_t id, const std::vector<float>& vector) override { data_.emplace_back(id, vector); }
  const std::vector<Point>& GetPoints() const override { return data_; }

 private:
  std::vector<Point> data_;
};

// status code -> the exception type the reference throws
inline void Check(vi_ctx* ctx, int rc)
{
  if (rc == VI_OK) return;
  const char* text = vi_last_error(ctx);
  switch (rc)
  {
    case VI_ERR_INVALID_ARG: throw std::invalid_argument(text);  // ArgumentException
    case VI_ERR_OVERFLOW: throw std::overflow_error(text);       // OverflowException, IndexBuilder.cs:99,104
    case VI_ERR_NOT_IMPLEMENTED: throw std::logic_error(text);   // NotImplementedException
    case VI_ERR_OOM: throw std::bad_alloc();
    default: throw std::runtime_error(text);
  }
}

struct CtxDeleter
{
  void operator()(vi_ctx* c) const { vi_destroy(c); }
};
using CtxPtr = std::unique_ptr<vi_ctx, CtxDeleter>;

inline CtxPtr MakeContext(int device = 0)
{
  vi_ctx* c = nullptr;
  if (vi_create(device, &c) != VI_OK) throw std::runtime_error("vi_create failed: no usable CUDA device (no CPU fallback)");
  return CtxPtr(c);
}

class IndexBuilder
{
 public:
  using StoreFactory = std::function<std::unique_ptr<IRangeStore>(int64_t rangeId, int64_t capacity)>;

  // Returns the (rangeId, RangeValue) rows Build yields (IndexBuilder.cs:92), breadth-first.  storeFactory is kept
  // for signature compatibility; child ranges stay on the device.
  static std::vector<std::pair<int64_t, RangeValue>> Build(const std::vector<Point>& points, const StoreFactory& = {},
                                                           int mode = VI_MODE_EXACT, int device = 0)
  {
    std::vector<std::pair<int64_t, RangeValue>> out;
    if (points.empty()) return out;  // IndexBuilder.cs:70-73
    CtxPtr ctx = MakeContext(device);
    const int32_t dims = (int32_t)points.front().second.size();
    Check(ctx.get(), vi_points_reserve(ctx.get(), (int64_t)points.size(), dims));
    std::vector<int64_t> ids;
    std::vector<float> rows;
    const size_t batch = 65536;
    for (size_t s = 0; s < points.size(); s += batch)
    {
      const size_t e = std::min(points.size(), s + batch);
      ids.clear();
      rows.clear();
      for (size_t i = s; i < e; ++i)
      {
        if ((int32_t)points[i].second.size() != dims) throw std::invalid_argument("Invalid length of vector.");
        ids.push_back(points[i].first);
        rows.insert(rows.end(), points[i].second.begin(), points[i].second.end());
      }
      Check(ctx.get(), vi_points_add(ctx.get(), ids.data(), rows.data(), (int64_t)(e - s), dims));
    }
    return Rows(ctx.get(), (int64_t)points.size(), mode);
  }

  // The real-data path of VectorIndex.MainTest (Program.cs:86-150): `/train` of an ANN-benchmarks HDF5 file, ids = row
  // indexes (Program.cs:252), streamed to the device by the library's native reader.
  static std::vector<std::pair<int64_t, RangeValue>> BuildFromHdf5(const char* fileName, const char* datasetName = "/train",
                                                                   int mode = VI_MODE_EXACT, int device = 0)
  {
    CtxPtr ctx = MakeContext(device);
    int64_t rows = 0, cols = 0;
    Check(ctx.get(), vi_hdf5_dataset_info(ctx.get(), fileName, datasetName, &rows, &cols, nullptr, nullptr, nullptr));
    if (rows == 0) return {};
    Check(ctx.get(), vi_points_reserve(ctx.get(), rows, (int32_t)cols));
    Check(ctx.get(), vi_points_add_hdf5(ctx.get(), fileName, datasetName, 0, -1, 0, nullptr, nullptr));
    return Rows(ctx.get(), rows, mode);
  }

 private:
  // vi_build_copy: build, then the rows -- the copy of the table's bulk overlaps the build's last kernel
  static std::vector<std::pair<int64_t, RangeValue>> Rows(vi_ctx* ctx, int64_t n, int mode)
  {
    const int64_t cap = 2 * n + n / 8 + 1024;  // 2n - 1 rows, plus one per one-sided split
    std::vector<int64_t> rid((size_t)cap), oid((size_t)cap);
    std::vector<int32_t> dim((size_t)cap);
    std::vector<float> mid((size_t)cap);
    int64_t k = 0;
    Check(ctx, vi_build_copy(ctx, mode, nullptr, rid.data(), dim.data(), mid.data(), oid.data(), cap, &k));
    std::vector<std::pair<int64_t, RangeValue>> out;
    out.reserve((size_t)k);
    for (int64_t i = 0; i < k; ++i) out.push_back({rid[i], RangeValue{dim[i], mid[i], oid[i]}});
    return out;
  }
};

// The consumer side of Program.cs:18-26: the rows a caller collected (or read back from the CSV of
// Program.cs:145-149) become a searchable index again (vi_ranges_load), and Search is dbo.Search (DDL.sql:234-295).
class RangeIndex
{
 public:
  RangeIndex(const std::vector<std::pair<int64_t, RangeValue>>& rows, int32_t dimensions, int device = 0)
      : ctx_(MakeContext(device)), dims_(dimensions)
  {
    std::vector<int64_t> rid, oid;
    std::vector<int32_t> dim;
    std::vector<float> mid;
    for (auto& [r, v] : rows)
    {
      rid.push_back(r);
      dim.push_back(v.Dimension);
      mid.push_back(v.Mid);
      oid.push_back(v.Id);
    }
    Check(ctx_.get(), vi_ranges_load(ctx_.get(), rid.data(), dim.data(), mid.data(), oid.data(), (int64_t)rid.size(), dims_));
  }

  // candidate ids of one query, traversal order (low branch first)
  std::vector<int64_t> Search(const std::vector<float>& vector, float proximity) const
  {
    if ((int32_t)vector.size() != dims_) throw std::invalid_argument("Invalid vector size.");  // MemoryVectorIndex.cs:254
    int64_t offsets[2] = {0, 0}, total = 0;
    Check(ctx_.get(), vi_search_begin(ctx_.get(), vector.data(), 1, dims_, proximity, &total));  // one walk: count ...
    std::vector<int64_t> ids((size_t)std::max<int64_t>(total, 1));
    Check(ctx_.get(), vi_search_fetch(ctx_.get(), offsets, ids.data(), (int64_t)ids.size()));    // ... then fetch
    ids.resize((size_t)total);
    return ids;
  }

 private:
  CtxPtr ctx_;
  int32_t dims_;
};
}  // namespace NesterovskyBros::VectorIndex
