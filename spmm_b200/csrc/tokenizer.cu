// Host-side WordPiece tokenizer over the 300-entry SMILES BPE vocabulary (reference SPMM_pretrain.py:19-20,
// SPMM_models.py:352: `self.tokenizer(text, padding='longest', truncation=True, max_length=100, return_tensors="pt")`
// with BertTokenizer(do_basic_tokenize=False) + WordpieceTokenizer(max_input_chars_per_word=250)).
// The reference tokenises in Python (greedy longest-match with substring slicing and dict lookups per candidate);
// at several thousand molecules per second per GPU that is the step's host bottleneck.  Here: one byte-trie for the
// word-initial pieces and one for the "##" continuation pieces, a whole batch per call, output written straight into
// the caller's (pinned) [n, ld] int64 buffers.  No CUDA in this file; it is part of the C-ABI library.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "spmm_b200.h"

namespace {

struct Trie {
  struct Node { int next[256]; int id; };
  std::vector<Node> nodes;
  Trie() { add_node(); }
  int add_node() {
    Node n;
    for (int i = 0; i < 256; ++i) n.next[i] = -1;
    n.id = -1;
    nodes.push_back(n);
    return (int)nodes.size() - 1;
  }
  void insert(const char* s, size_t len, int id) {
    int cur = 0;
    for (size_t i = 0; i < len; ++i) {
      const unsigned char c = (unsigned char)s[i];
      if (nodes[cur].next[c] < 0) { const int nn = add_node(); nodes[cur].next[c] = nn; }
      cur = nodes[cur].next[c];
    }
    if (nodes[cur].id < 0) nodes[cur].id = id;   // first occurrence wins, like a Python dict built from the vocab file order
  }
  // longest vocabulary entry that is a prefix of s[0..len): returns its id and sets *mlen, or -1
  int longest(const char* s, size_t len, size_t* mlen) const {
    int cur = 0, best = -1;
    for (size_t i = 0; i < len; ++i) {
      cur = nodes[cur].next[(unsigned char)s[i]];
      if (cur < 0) break;
      if (nodes[cur].id >= 0) { best = nodes[cur].id; *mlen = i + 1; }
    }
    return best;
  }
};

struct WordPiece {
  Trie initial, cont;
  int unk_id, max_chars;
};

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// number of unicode characters of a UTF-8 byte range (the reference limit counts characters)
inline size_t utf8_chars(const char* s, size_t len) {
  size_t n = 0;
  for (size_t i = 0; i < len; ++i) n += (((unsigned char)s[i]) & 0xC0) != 0x80;
  return n;
}

}  // namespace

extern "C" void* spmm_wordpiece_create(const char* const* tokens, int n_tokens, int unk_id, int max_input_chars_per_word) {
  if (!tokens || n_tokens <= 0 || unk_id < 0 || unk_id >= n_tokens) return nullptr;
  WordPiece* wp = new WordPiece();
  wp->unk_id = unk_id;
  wp->max_chars = max_input_chars_per_word;
  for (int i = 0; i < n_tokens; ++i) {
    const char* t = tokens[i];
    const size_t len = strlen(t);
    if (len > 2 && t[0] == '#' && t[1] == '#') wp->cont.insert(t + 2, len - 2, i);
    wp->initial.insert(t, len, i);   // "##x" is also a legal word-initial piece for a word that literally starts with "##x"
  }
  return wp;
}

extern "C" void spmm_wordpiece_destroy(void* handle) { delete static_cast<WordPiece*>(handle); }

// Encodes n texts: ids = [cls] + pieces (truncated to max_length - 2) + [sep], right-padded with pad_id to `width` =
// min(max_length, longest sequence of the batch) (padding='longest').  ids_out / mask_out are [n][ld] int64 (ld >= width;
// columns >= width are left untouched).  Returns width, or a negative value on bad arguments (-2: ld too small).
extern "C" int spmm_wordpiece_encode_batch(void* handle, const char* const* texts, int n, int max_length, int cls_id,
                                           int sep_id, int pad_id, int64_t* ids_out, int64_t* mask_out, int ld) {
  if (!handle || !texts || n <= 0 || max_length < 2 || !ids_out || !mask_out) return -1;
  const WordPiece* wp = static_cast<const WordPiece*>(handle);
  std::vector<std::vector<int> > all((size_t)n);
  int width = 0;
  std::vector<int> word;
  for (int t = 0; t < n; ++t) {
    std::vector<int>& out = all[(size_t)t];
    out.push_back(cls_id);
    const char* s = texts[t];
    const size_t len = strlen(s);
    size_t i = 0;
    while (i < len) {                                   // whitespace_tokenize
      while (i < len && is_space((unsigned char)s[i])) ++i;
      size_t j = i;
      while (j < len && !is_space((unsigned char)s[j])) ++j;
      if (j == i) break;
      const char* w = s + i;
      const size_t wl = j - i;
      if ((int)utf8_chars(w, wl) > wp->max_chars) {
        out.push_back(wp->unk_id);
      } else {
        word.clear();
        size_t start = 0;
        bool bad = false;
        while (start < wl) {                            // greedy longest-match-first
          size_t ml = 0;
          const int id = (start == 0 ? wp->initial : wp->cont).longest(w + start, wl - start, &ml);
          if (id < 0) { bad = true; break; }
          word.push_back(id);
          start += ml;
        }
        if (bad) out.push_back(wp->unk_id);             // any miss turns the whole word into [UNK]
        else out.insert(out.end(), word.begin(), word.end());
      }
      i = j;
    }
    if ((int)out.size() > max_length - 1) out.resize((size_t)max_length - 1);   // truncation keeps [cls] + first max_length-2 pieces
    out.push_back(sep_id);
    if ((int)out.size() > width) width = (int)out.size();
  }
  if (width > ld) return -2;
  for (int t = 0; t < n; ++t) {
    const std::vector<int>& o = all[(size_t)t];
    int64_t* ids = ids_out + (size_t)t * ld;
    int64_t* mk = mask_out + (size_t)t * ld;
    int c = 0;
    for (; c < (int)o.size(); ++c) { ids[c] = o[(size_t)c]; mk[c] = 1; }
    for (; c < width; ++c) { ids[c] = pad_id; mk[c] = 0; }
  }
  return width;
}
