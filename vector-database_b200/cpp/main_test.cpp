// main_test.cpp -- the shape of VectorIndex.MainTest/Program.cs:9-67 over the C++ mirror: random set + the crafted
// one-hot set through IndexBuilder::Build with a MemoryRangeStore factory.  Prints "rangeId,Dimension,Mid(bits),Id"
// rows (Program.cs:80 CSV shape) so a test can diff them against the oracle.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "vector_index.hpp"

using namespace NesterovskyBros::VectorIndex;

static void print_rows(const std::vector<std::pair<int64_t, RangeValue>>& index)
{
  for (auto& [rangeId, range] : index)
  {
    uint32_t bits;
    memcpy(&bits, &range.Mid, 4);
    printf("%lld,%d,%u,%lld\n", (long long)rangeId, range.Dimension, bits, (long long)range.Id);
  }
}

int main(int argc, char** argv)
{
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  if (argc > 3 && strcmp(argv[2], "hdf5") == 0)
  {
    // main_test <mode> hdf5 <file> [dataset]: the deep-image path of Program.cs:86-150
    try
    {
      print_rows(IndexBuilder::BuildFromHdf5(argv[3], argc > 4 ? argv[4] : "/train", mode));
    }
    catch (const std::exception& e)
    {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
    return 0;
  }
  const int dimensions = argc > 2 ? atoi(argv[2]) : 1536;
  std::vector<Point> input;  // Program.cs:54-66
  for (int64_t i = 0; i < dimensions; ++i)
  {
    std::vector<float> v(dimensions, 0.0f);
    v[i] = 1.0f;
    input.emplace_back(i, std::move(v));
  }
  try
  {
    auto index = IndexBuilder::Build(input, [](int64_t, int64_t) { return std::make_unique<MemoryRangeStore>(); }, mode);
    print_rows(index);
    // the rows alone are a searchable index (vi_ranges_load): a point lookup at proximity 0 must return the point
    RangeIndex lookup(index, dimensions);
    for (int64_t i = 0; i < dimensions && i < 16; ++i)
    {
      const std::vector<int64_t> ids = lookup.Search(input[(size_t)i].second, 0.0f);
      if (std::find(ids.begin(), ids.end(), i) == ids.end())
      {
        fprintf(stderr, "error: point %lld not found by the imported table\n", (long long)i);
        return 2;
      }
    }
  }
  catch (const std::exception& e)
  {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
