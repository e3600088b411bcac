// Index builders for the memory-mapped token datasets (pybind11 module `fast_index_map_helpers`).
//
// Output contracts are those of the reference helper (ppfleetx/data/data_tools/cpp/fast_index_map_helpers.cpp:
// build_sample_idx :92, build_mapping :194/:431, build_blocks_mapping :455/:671, build_blending_indices :32) so
// that cached `*_idx.npy` files stay interchangeable — same arrays, bit for bit, same RNG consumption order
// (mt19937(seed) for target lengths, mt19937_64(seed + 1) Fisher-Yates at the end).  The implementation is
// a single pass into growable vectors driven by one generic sentence-span packer.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace py = pybind11;

namespace {

constexpr int32_t kLongSentence = 512;   // documents containing a longer sentence are skipped

template <typename T>
py::array take_ownership(std::vector<T>* vec, std::vector<py::ssize_t> shape) {
  py::capsule owner(vec, [](void* p) { delete reinterpret_cast<std::vector<T>*>(p); });
  std::vector<py::ssize_t> strides(shape.size());
  py::ssize_t s = sizeof(T);
  for (int i = (int)shape.size() - 1; i >= 0; --i) { strides[i] = s; s *= shape[i]; }
  return py::array(py::dtype::of<T>(), shape, strides, vec->data(), owner);
}

// ---------------------------------------------------------------------------------- GPT sample index
// sample_idx[i] = (index into doc_idx, token offset inside that document) of the first token of sample i;
// consecutive samples overlap by one token (labels are inputs shifted by one).
py::array build_sample_idx(const py::array_t<int64_t>& sizes_, const py::array_t<int64_t>& doc_idx_, int32_t seq_length,
                           int32_t num_epochs, int64_t tokens_per_epoch) {
  if (seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) throw std::invalid_argument("build_sample_idx: bad arguments");
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = doc_idx_.unchecked<1>();
  const int64_t num_samples = (static_cast<int64_t>(num_epochs) * tokens_per_epoch - 1) / seq_length;
  auto* out = new std::vector<int64_t>(2 * (num_samples + 1));
  int64_t cursor_doc = 0, cursor_off = 0;
  (*out)[0] = 0; (*out)[1] = 0;
  for (int64_t s = 1; s <= num_samples; ++s) {
    int64_t need = seq_length + 1;           // seq_length inputs + 1 label lookahead
    while (need > 0) {
      const int64_t avail = sizes[doc_idx[cursor_doc]] - cursor_off;
      if (avail >= need) {
        cursor_off += need - 1;              // keep the last token: it is the first input of the next sample
        need = 0;
      } else {
        need -= avail;
        ++cursor_doc;
        cursor_off = 0;
      }
    }
    (*out)[2 * s] = cursor_doc;
    (*out)[2 * s + 1] = cursor_off;
  }
  return take_ownership(out, {num_samples + 1, 2});
}

// ---------------------------------------------------------------------------------- sentence-span packing
struct DocView {
  int64_t first, last;    // sentence index range [first, last)
  int32_t doc;
};

// Walks the sentences of one document, closing a span whenever `close(seq_len, num_sent, remaining)` says so or
// the document ends; `emit(start, end)` is called per span.  Returns nothing; caller owns RNG draws via callbacks.
template <typename Close, typename Emit>
inline void pack_document(const DocView& d, const py::detail::unchecked_reference<int32_t, 1>& sizes, Close close, Emit emit) {
  int64_t start = d.first;
  int32_t seq_len = 0, num_sent = 0;
  int64_t remaining = d.last - d.first;
  for (int64_t s = d.first; s < d.last; ++s) {
    seq_len += sizes[s];
    ++num_sent;
    --remaining;
    if (remaining == 0 || close(seq_len, num_sent, remaining)) {
      emit(start, s + 1);
      start = s + 1;
      seq_len = 0;
      num_sent = 0;
    }
  }
}

inline bool has_long_sentence(const DocView& d, const py::detail::unchecked_reference<int32_t, 1>& sizes) {
  for (int64_t s = d.first; s < d.last; ++s)
    if (sizes[s] > kLongSentence) return true;
  return false;
}

template <typename T>
void shuffle_rows(std::vector<T>& rows, int width, int32_t seed) {
  const int64_t n = (int64_t)rows.size() / width;
  std::mt19937_64 gen(static_cast<uint64_t>(seed + 1));
  for (int64_t i = n - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(gen() % static_cast<uint64_t>(i + 1));
    for (int c = 0; c < width; ++c) std::swap(rows[i * width + c], rows[j * width + c]);
  }
}

inline int32_t draw_target_len(int32_t short_ratio, int32_t max_len, std::mt19937& gen) {
  if (short_ratio == 0) return max_len;
  const auto r = gen();
  return (r % short_ratio) == 0 ? static_cast<int32_t>(2 + r % (max_len - 1)) : max_len;
}

template <typename T>
py::array build_mapping_t(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, int32_t num_epochs,
                          uint64_t max_num_samples, int32_t max_seq_length, double short_seq_prob, int32_t seed, bool verbose,
                          int32_t min_num_sent) {
  if (num_epochs <= 0 || max_seq_length <= 1 || short_seq_prob < 0 || short_seq_prob > 1 || seed <= 0)
    throw std::invalid_argument("build_mapping: bad arguments");
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  const int32_t short_ratio = short_seq_prob > 0 ? static_cast<int32_t>(std::round(1.0 / short_seq_prob)) : 0;
  auto* rows = new std::vector<T>();
  std::mt19937 gen(static_cast<uint32_t>(seed));
  uint64_t count = 0;
  for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
    if (count >= max_num_samples) break;
    if (epoch > 0 && count == 0) {
      delete rows;
      throw std::invalid_argument("Invalid dataset! the document should be with more than " + std::to_string(min_num_sent) + " scentences.");
    }
    for (int32_t doc = 0; doc < docs.shape(0) - 1; ++doc) {
      const DocView d{docs[doc], docs[doc + 1], doc};
      const int64_t n_sent = d.last - d.first;
      if (n_sent < min_num_sent) continue;
      if (n_sent > 1 && has_long_sentence(d, sizes)) continue;
      int32_t target = draw_target_len(short_ratio, max_seq_length, gen);
      pack_document(
          d, sizes,
          [&](int32_t seq_len, int32_t num_sent, int64_t remaining) {
            return seq_len >= target && remaining > 1 && num_sent >= min_num_sent;
          },
          [&](int64_t a, int64_t b) {
            rows->push_back(static_cast<T>(a));
            rows->push_back(static_cast<T>(b));
            rows->push_back(static_cast<T>(target));
            ++count;
            target = draw_target_len(short_ratio, max_seq_length, gen);
          });
    }
  }
  shuffle_rows(*rows, 3, seed);
  (void)verbose;
  return take_ownership(rows, {(py::ssize_t)count, 3});
}

py::array build_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes, int num_epochs, uint64_t max_num_samples,
                        int max_seq_length, double short_seq_prob, int seed, bool verbose, int32_t min_num_sent) {
  if ((uint64_t)sizes.size() > std::numeric_limits<uint32_t>::max())
    return build_mapping_t<uint64_t>(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose, min_num_sent);
  return build_mapping_t<uint32_t>(docs, sizes, num_epochs, max_num_samples, max_seq_length, short_seq_prob, seed, verbose, min_num_sent);
}

template <typename T>
py::array build_blocks_mapping_t(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_,
                                 const py::array_t<int32_t>& titles_sizes_, int32_t num_epochs, uint64_t max_num_samples,
                                 int32_t max_seq_length, int32_t seed, bool verbose, bool use_one_sent_blocks) {
  if (num_epochs <= 0 || max_seq_length <= 1 || seed <= 0) throw std::invalid_argument("build_blocks_mapping: bad arguments");
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  auto titles = titles_sizes_.unchecked<1>();
  const int32_t min_num_sent = use_one_sent_blocks ? 1 : 2;
  auto* rows = new std::vector<T>();
  uint64_t count = 0;
  for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
    if (count >= max_num_samples) break;
    int32_t block_id = 0;
    for (int32_t doc = 0; doc < docs.shape(0) - 1; ++doc) {
      const DocView d{docs[doc], docs[doc + 1], doc};
      const int64_t n_sent = d.last - d.first;
      if (n_sent < min_num_sent || has_long_sentence(d, sizes)) continue;
      const int32_t target = max_seq_length - titles[doc];
      pack_document(
          d, sizes,
          [&](int32_t seq_len, int32_t num_sent, int64_t remaining) {
            return seq_len >= target && remaining >= min_num_sent && num_sent >= min_num_sent;
          },
          [&](int64_t a, int64_t b) {
            rows->push_back(static_cast<T>(a));
            rows->push_back(static_cast<T>(b));
            rows->push_back(static_cast<T>(doc));
            rows->push_back(static_cast<T>(block_id));
            ++count;
            ++block_id;
          });
    }
  }
  shuffle_rows(*rows, 4, seed);
  (void)verbose;
  return take_ownership(rows, {(py::ssize_t)count, 4});
}

py::array build_blocks_mapping(const py::array_t<int64_t>& docs, const py::array_t<int32_t>& sizes, const py::array_t<int32_t>& titles,
                               int num_epochs, uint64_t max_num_samples, int max_seq_length, int seed, bool verbose,
                               bool use_one_sent_blocks) {
  if ((uint64_t)sizes.size() > std::numeric_limits<uint32_t>::max())
    return build_blocks_mapping_t<uint64_t>(docs, sizes, titles, num_epochs, max_num_samples, max_seq_length, seed, verbose, use_one_sent_blocks);
  return build_blocks_mapping_t<uint32_t>(docs, sizes, titles, num_epochs, max_num_samples, max_seq_length, seed, verbose, use_one_sent_blocks);
}

// ---------------------------------------------------------------------------------- weighted dataset blending
// Greedy largest-deficit assignment: at step i pick the dataset whose achieved count lags its target weight * i most.
void build_blending_indices(py::array_t<uint8_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index,
                            const py::array_t<double>& weights, int32_t num_datasets, int64_t size, bool verbose) {
  auto di = dataset_index.mutable_unchecked<1>();
  auto si = dataset_sample_index.mutable_unchecked<1>();
  auto w = weights.unchecked<1>();
  std::vector<int64_t> taken(num_datasets, 0);
  for (int64_t i = 0; i < size; ++i) {
    const double t = std::max(static_cast<double>(i), 1.0);
    int64_t best = 0;
    double best_err = w[0] * t - static_cast<double>(taken[0]);
    for (int64_t k = 1; k < num_datasets; ++k) {
      const double err = w[k] * t - static_cast<double>(taken[k]);
      if (err > best_err) { best_err = err; best = k; }
    }
    di[i] = static_cast<uint8_t>(best);
    si[i] = taken[best]++;
  }
  (void)verbose;
}

}  // namespace

PYBIND11_MODULE(fast_index_map_helpers, m) {
  m.doc() = "dataset index builders (sample index, sentence-pair spans, ICT blocks, blending)";
  m.def("build_sample_idx", &build_sample_idx);
  m.def("build_mapping", &build_mapping);
  m.def("build_blocks_mapping", &build_blocks_mapping);
  m.def("build_blending_indices", &build_blending_indices);
}
