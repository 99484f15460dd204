// De Bruijn graph of the local realigner on the host (native): the candidate haplotypes of one window.
//
// Restates deepvariant/realigner/debruijn_graph.cc: AddEdgesForReference / AddEdgesForRead (:250-300: k-mers of the reference, and
// of every run of usable read bases - canonical, base quality >= min_base_quality, reads with mapping quality >= min_mapq), cycle
// detection (:175-222), Prune (:302-345: edges that are on the reference or seen >= min_edge_weight times; vertices reachable from
// the source AND reaching the sink), CandidateHaplotypes (:347-391: every source -> sink / dead-end path, none when more than
// max_num_paths are in flight), Build (:224-248: the smallest k for which neither the reference nor the graph has a cycle).
// The Python restatement (deepvariant_b200/realigner.py, DeBruijnGraph) stays as the cross-check of the tests; the host profile of
// the product CLI put 0.84 s of every 10 kb here.  Host code only - compiled into libdvb.so with the rest.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "dvb_common.h"

namespace {

struct Edge { int to; int weight; bool is_ref; };

struct Graph {
  int k = 0;
  std::unordered_map<std::string_view, int> id;
  std::vector<std::string_view> kmer;
  std::vector<std::vector<Edge>> out;

  int Vertex(std::string_view s) {
    auto it = id.find(s);
    if (it != id.end()) return it->second;
    const int v = (int)kmer.size();
    id.emplace(s, v);
    kmer.push_back(s);
    out.emplace_back();
    return v;
  }

  // bases[start .. end + k): the k-mers at start .. end and the edges between neighbours
  void Add(const char* bases, int64_t start, int64_t end, bool is_ref) {
    if (end <= 0) return;
    int prev = Vertex(std::string_view(bases + start, (size_t)k));
    for (int64_t i = start + 1; i <= end; ++i) {
      const int cur = Vertex(std::string_view(bases + i, (size_t)k));
      std::vector<Edge>& es = out[(size_t)prev];
      Edge* e = nullptr;
      for (Edge& x : es)
        if (x.to == cur) { e = &x; break; }
      if (!e) { es.push_back(Edge{cur, 0, false}); e = &es.back(); }
      ++e->weight;
      e->is_ref = e->is_ref || is_ref;
      prev = cur;
    }
  }

  bool HasCycle() const {
    const int n = (int)kmer.size();
    std::vector<uint8_t> color((size_t)n, 0);
    std::vector<std::pair<int, size_t>> stack;
    for (int root = 0; root < n; ++root) {
      if (color[(size_t)root]) continue;
      color[(size_t)root] = 1;
      stack.emplace_back(root, 0);
      while (!stack.empty()) {
        const int v = stack.back().first;
        size_t& next = stack.back().second;
        if (next < out[(size_t)v].size()) {
          const int w = out[(size_t)v][next++].to;
          if (color[(size_t)w] == 1) return true;
          if (color[(size_t)w] == 0) { color[(size_t)w] = 1; stack.emplace_back(w, 0); }
        } else {
          color[(size_t)v] = 2;
          stack.pop_back();
        }
      }
    }
    return false;
  }

  void Prune(int source, int sink, int min_edge_weight) {
    const int n = (int)kmer.size();
    for (auto& es : out)
      es.erase(std::remove_if(es.begin(), es.end(), [&](const Edge& e) { return !e.is_ref && e.weight < min_edge_weight; }), es.end());
    std::vector<std::vector<int>> rev((size_t)n);
    for (int v = 0; v < n; ++v)
      for (const Edge& e : out[(size_t)v]) rev[(size_t)e.to].push_back(v);
    auto reach = [&](int root, auto&& next) {
      std::vector<uint8_t> seen((size_t)n, 0);
      std::vector<int> st{root};
      seen[(size_t)root] = 1;
      while (!st.empty()) {
        const int v = st.back();
        st.pop_back();
        next(v, [&](int w) { if (!seen[(size_t)w]) { seen[(size_t)w] = 1; st.push_back(w); } });
      }
      return seen;
    };
    const std::vector<uint8_t> fwd = reach(source, [&](int v, auto&& visit) { for (const Edge& e : out[(size_t)v]) visit(e.to); });
    const std::vector<uint8_t> bwd = reach(sink, [&](int v, auto&& visit) { for (int w : rev[(size_t)v]) visit(w); });
    for (int v = 0; v < n; ++v) {
      if (!(fwd[(size_t)v] && bwd[(size_t)v])) { out[(size_t)v].clear(); continue; }
      auto& es = out[(size_t)v];
      es.erase(std::remove_if(es.begin(), es.end(), [&](const Edge& e) { return !(fwd[(size_t)e.to] && bwd[(size_t)e.to]); }), es.end());
    }
    kept_ = std::vector<uint8_t>((size_t)n, 0);
    for (int v = 0; v < n; ++v) kept_[(size_t)v] = fwd[(size_t)v] && bwd[(size_t)v];
  }

  std::vector<uint8_t> kept_;
};

bool Usable(uint8_t base, uint8_t qual, int min_base_quality) {
  return (base == 'A' || base == 'C' || base == 'G' || base == 'T') && qual >= min_base_quality;
}

}  // namespace

extern "C" {

// Candidate haplotypes of one window, sorted, '\n'-separated, into `out` (cap bytes).  Returns the number of bytes the text needs
// (written only when it fits; 0 = no graph: the caller keeps the reference alone), or a negative DvbStatus.
// reads: upper- or lower-case bases and qualities concatenated, read i at [read_begin[i], read_begin[i + 1]); mapq per read.
int64_t dvb_dbg_candidate_haplotypes(const char* ref, int64_t ref_len, const char* bases, const uint8_t* quals, const int64_t* read_begin,
                                     const int32_t* mapq, int32_t n_reads, int32_t min_k, int32_t max_k, int32_t step_k, int32_t min_mapq,
                                     int32_t min_base_quality, int32_t min_edge_weight, int32_t max_num_paths, char* out, int64_t cap,
                                     int32_t* k_used) {
  if (!ref || ref_len < 0 || n_reads < 0 || (n_reads > 0 && (!bases || !quals || !read_begin || !mapq)) || step_k < 1 || min_k < 1 || (cap > 0 && !out))
    return -(int64_t)dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_dbg_candidate_haplotypes: bad arguments");
  if (k_used) *k_used = 0;
  // upper-case copies (aligned_sequence.upper() in the Python form)
  std::string up;
  if (n_reads > 0) {
    up.assign(bases + read_begin[0], (size_t)(read_begin[n_reads] - read_begin[0]));
    for (char& c : up)
      if (c >= 'a' && c <= 'z') c = (char)(c - 32);
  }
  const int64_t base0 = n_reads > 0 ? read_begin[0] : 0;
  const int hi_k = (int)std::min<int64_t>(max_k, ref_len - 1);
  int first_k = -1;
  for (int k = min_k; k <= hi_k; k += step_k) {       // the reference alone must not repeat a k-mer
    std::unordered_map<std::string_view, int> seen;
    bool unique = true;
    for (int64_t i = 0; i + k <= ref_len && unique; ++i) unique = seen.emplace(std::string_view(ref + i, (size_t)k), 0).second;
    if (unique) { first_k = k; break; }
  }
  if (first_k < 0) return 0;
  for (int k = first_k; k <= hi_k; k += step_k) {
    Graph g;
    g.k = k;
    g.Add(ref, 0, ref_len - k, true);
    for (int r = 0; r < n_reads; ++r) {
      if (mapq[r] < min_mapq) continue;
      const char* b = up.data() + (read_begin[r] - base0);
      const uint8_t* q = quals + read_begin[r];
      const int64_t n = read_begin[r + 1] - read_begin[r];
      const int64_t stop = n - k;
      int64_t i = 0;
      while (i < stop) {
        int64_t bad = n;
        for (int64_t j = i; j < n; ++j)
          if (!Usable((uint8_t)b[j], q[j], min_base_quality)) { bad = j; break; }
        g.Add(b, i, bad - k, false);
        i = bad + 1;
      }
    }
    if (g.HasCycle()) continue;
    const int source = g.id.at(std::string_view(ref, (size_t)k));
    const int sink = g.id.at(std::string_view(ref + ref_len - k, (size_t)k));
    g.Prune(source, sink, min_edge_weight);
    if (k_used) *k_used = k;
    // every path from the source to the sink or a dead end, breadth first; too many in flight -> none
    std::vector<std::vector<int>> terminated, queue{{source}};
    size_t head = 0;
    bool overflow = false;
    while (head < queue.size()) {
      if (terminated.size() + (queue.size() - head) > (size_t)max_num_paths) { overflow = true; break; }
      const std::vector<int> path = queue[head++];
      for (const Edge& e : g.out[(size_t)path.back()]) {
        std::vector<int> ext = path;
        ext.push_back(e.to);
        if (e.to == sink || g.out[(size_t)e.to].empty()) terminated.push_back(std::move(ext));
        else queue.push_back(std::move(ext));
      }
      if (head > 4096 && head * 2 > queue.size()) {   // drop the consumed prefix now and then
        queue.erase(queue.begin(), queue.begin() + (long)head);
        head = 0;
      }
    }
    std::vector<std::string> haps;
    if (!overflow) {
      for (const std::vector<int>& p : terminated) {
        std::string h;
        h.reserve(p.size() + (size_t)k);
        for (int v : p) h.push_back(g.kmer[(size_t)v][0]);
        h.append(g.kmer[(size_t)p.back()].substr(1));
        haps.push_back(std::move(h));
      }
      std::sort(haps.begin(), haps.end());
    }
    int64_t need = 1;      // at least one byte so that "a graph with no haplotype" (1) differs from "no graph" (0)
    for (const std::string& h : haps) need += (int64_t)h.size() + 1;
    if (need <= cap) {
      char* w = out;
      for (const std::string& h : haps) { memcpy(w, h.data(), h.size()); w += h.size(); *w++ = '\n'; }
      *w = '\0';
    }
    return need;
  }
  return 0;
}

}  // extern "C"
