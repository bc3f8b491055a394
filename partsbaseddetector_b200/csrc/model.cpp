// model.cpp -- see model.hpp.  A small purpose-built reader for the `opencv_storage`
// XML subset the reference's models use (scalars, opencv-matrix with dt=d, flat numeric
// sequences, sequences of sequences, nested maps) and a writer producing the same schema
// as FileStorageModel::serialize (reference src/FileStorageModel.cpp:42-94).
#include "model.hpp"

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <memory>
#include <stdexcept>

namespace pbd {
namespace {

struct XNode {
  std::string name;
  std::string type_id;             // attribute type_id, if any
  const char* tb = nullptr;        // text range (only meaningful for leaves)
  const char* te = nullptr;
  std::vector<std::unique_ptr<XNode>> kids;
  const XNode* child(const char* n) const {
    for (auto& k : kids) if (k->name == n) return k.get();
    return nullptr;
  }
};

struct XParser {
  const char* p;
  const char* end;
  explicit XParser(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
  [[noreturn]] void fail(const char* what) { throw FormatError(std::string("XML: ") + what); }
  void skip_ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p; }
  void skip_misc() {   // whitespace, <?...?>, <!-- ... -->
    for (;;) {
      skip_ws();
      if (p + 1 < end && p[0] == '<' && p[1] == '?') { const char* q = strstr(p, "?>"); if (!q) fail("unterminated <?"); p = q + 2; }
      else if (p + 3 < end && !strncmp(p, "<!--", 4)) { const char* q = strstr(p, "-->"); if (!q) fail("unterminated comment"); p = q + 3; }
      else return;
    }
  }
  std::unique_ptr<XNode> element() {
    if (p >= end || *p != '<') fail("expected '<'");
    ++p;
    auto n = std::make_unique<XNode>();
    const char* s = p;
    while (p < end && *p != ' ' && *p != '>' && *p != '/' && *p != '\n' && *p != '\t') ++p;
    n->name.assign(s, p);
    // attributes
    for (;;) {
      skip_ws();
      if (p >= end) fail("eof in tag");
      if (*p == '>') { ++p; break; }
      if (*p == '/' && p + 1 < end && p[1] == '>') { p += 2; return n; }
      const char* as = p;
      while (p < end && *p != '=' && *p != '>' && *p != ' ') ++p;
      std::string an(as, p);
      if (p < end && *p == '=') {
        ++p;
        if (p >= end || (*p != '"' && *p != '\'')) fail("bad attribute");
        const char q = *p++;
        const char* vs = p;
        while (p < end && *p != q) ++p;
        if (p >= end) fail("unterminated attribute");
        if (an == "type_id") n->type_id.assign(vs, p);
        ++p;
      }
    }
    // content
    n->tb = p;
    for (;;) {
      const char* lt = (const char*)memchr(p, '<', end - p);
      if (!lt) fail("eof in element");
      if (lt + 1 < end && lt[1] == '/') {
        if (n->kids.empty()) n->te = lt; else n->te = n->tb;
        p = lt + 2;
        const char* cs = p;
        while (p < end && *p != '>') ++p;
        if (std::string(cs, p) != n->name) fail("mismatched closing tag");
        ++p;
        return n;
      }
      if (lt + 3 < end && !strncmp(lt, "<!--", 4)) { const char* q = strstr(lt, "-->"); if (!q) fail("unterminated comment"); p = q + 3; continue; }
      p = lt;
      n->kids.push_back(element());
    }
  }
};

template <typename F>
void for_each_number(const XNode* n, F f) {
  if (!n) return;
  const char* p = n->tb;
  const char* e = n->te;
  std::string buf(p, e);             // strtod needs NUL termination
  const char* c = buf.c_str();
  for (;;) {
    while (*c == ' ' || *c == '\n' || *c == '\r' || *c == '\t') ++c;
    if (!*c) break;
    char* q = nullptr;
    double v;
    // cv::FileStorage writes .Inf / -.Inf / .NaN; the tokens are matched in full so that a truncated one cannot run past the buffer
    if (!strncasecmp(c, ".inf", 4)) { v = INFINITY; q = const_cast<char*>(c) + 4; }
    else if (!strncasecmp(c, ".nan", 4)) { v = NAN; q = const_cast<char*>(c) + 4; }
    else if (!strncasecmp(c, "-.inf", 5)) { v = -INFINITY; q = const_cast<char*>(c) + 5; }
    else v = strtod(c, &q);
    if (q == c) throw FormatError("XML: bad number in <" + n->name + ">");
    f(v);
    c = q;
  }
}
double scalar(const XNode* root, const char* name) {
  const XNode* n = root->child(name);
  if (!n) throw FormatError(std::string("XML: missing <") + name + ">");
  double v = 0; int cnt = 0;
  for_each_number(n, [&](double x) { if (!cnt++) v = x; });
  if (!cnt) throw FormatError(std::string("XML: empty <") + name + ">");
  return v;
}
std::string trimmed(const XNode* n) {
  std::string s(n->tb, n->te);
  size_t a = s.find_first_not_of(" \n\r\t\""), b = s.find_last_not_of(" \n\r\t\"");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

std::string slurp(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw IoError("cannot open '" + path + "': " + strerror(errno));
  std::string s;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n > 0) { s.resize((size_t)n); if (fread(&s[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); throw IoError("short read on '" + path + "'"); } }
  fclose(f);
  return s;
}

}  // namespace

void Model::validate() const {
  auto bad = [](const std::string& s) { throw FormatError("model: " + s); };
  if (interval <= 0 || sbin <= 0 || flen <= 0 || norient <= 0 || norient % 2) bad("bad header");
  if (flen != 3 * norient / 2 + 5) bad("flen must be 3*norient/2+5 (HOG layout, reference src/HOGFeatures.cpp:304-338)");
  if (filters.empty() || comps.empty()) bad("no filters / components");
  if (anchors.size() != (size_t)ndefs() * 2) bad("anchors/defs size mismatch");
  for (int i = 0; i < nfilters(); ++i)
    if (frows[i] <= 0 || fkw[i] <= 0 || filters[i].size() != (size_t)frows[i] * fkw[i] * flen) bad("bad filter shape");
  for (auto& c : comps) {
    if (c.empty()) bad("empty component");
    for (size_t p = 0; p < c.size(); ++p) {
      const Part& P = c[p];
      if (P.filterid.empty()) bad("part without filters");
      if (p == 0 ? P.parentid >= 0 : (P.parentid < 0 || P.parentid >= (int)p)) bad("parts must be ordered root-first (parentid < self)");
      for (int f : P.filterid) if (f < 0 || f >= nfilters()) bad("filterid out of range");
      if (P.biasid.empty()) bad("part without biasid");
      if (p == 0) { if (P.biasid[0] < 0 || P.biasid[0] >= (int)biasw.size()) bad("root biasid out of range"); continue; }
      const int pn = (int)c[P.parentid].filterid.size();
      if (P.biasid.size() < P.filterid.size() || P.defid.size() < P.filterid.size()) bad("biasid/defid shorter than filterid");
      for (size_t mm = 0; mm < P.filterid.size(); ++mm) {
        if (P.biasid[mm] < 0 || P.biasid[mm] + pn > (int)biasw.size()) bad("biasid out of range");
        if (P.defid[mm] < 0 || P.defid[mm] >= ndefs()) bad("defid out of range");
        if (!(defs[P.defid[mm] * 4] > 0.f) || !(defs[P.defid[mm] * 4 + 2] > 0.f)) bad("quadratic deformation weights must be > 0");
      }
    }
  }
}

// the schema of FileStorageModel::serialize (reference src/FileStorageModel.cpp:42-94), from a parsed XML or YAML tree
static void model_from_tree(const XNode* root, Model& m);

void load_xml(const std::string& path, Model& m) {
  const std::string text = slurp(path);
  XParser xp(text);
  xp.skip_misc();
  std::unique_ptr<XNode> root = xp.element();
  if (root->name != "opencv_storage") throw FormatError("XML: root element is not <opencv_storage>");
  model_from_tree(root.get(), m);
}

static void model_from_tree(const XNode* root, Model& m) {
  m = Model();
  if (const XNode* n = root->child("name")) m.name = trimmed(n);
  m.interval = (int)scalar(root, "interval");
  m.thresh = (float)scalar(root, "thresh");
  m.sbin = (int)scalar(root, "sbin");
  m.norient = (int)scalar(root, "norient");
  m.flen = (int)scalar(root, "flen");
  const XNode* fw = root->child("filtersw");
  if (!fw) throw FormatError("XML: missing <filtersw>");
  for (auto& k : fw->kids) {
    const int rows = (int)scalar(k.get(), "rows"), cols = (int)scalar(k.get(), "cols");
    const XNode* dt = k->child("dt");
    const std::string dts = dt ? trimmed(dt) : "d";
    if (dts != "d" && dts != "f") throw FormatError("XML: filter matrix dt must be d or f");
    if (m.flen <= 0 || cols % m.flen) throw FormatError("XML: filter cols not a multiple of flen");
    std::vector<double> v;
    v.reserve((size_t)rows * cols);
    for_each_number(k->child("data"), [&](double x) { v.push_back(dts == "f" ? (double)(float)x : x); });
    if (v.size() != (size_t)rows * cols) throw FormatError("XML: filter data size mismatch");
    m.frows.push_back(rows);
    m.fkw.push_back(cols / m.flen);
    m.filters.push_back(std::move(v));
  }
  for_each_number(root->child("biasw"), [&](double x) { m.biasw.push_back((float)x); });
  for_each_number(root->child("anchors"), [&](double x) { m.anchors.push_back((int)x); });
  if (const XNode* defs = root->child("defs"))
    for (auto& k : defs->kids) {
      int cnt = 0;
      for_each_number(k.get(), [&](double x) { m.defs.push_back((float)x); ++cnt; });
      if (cnt != 4) throw FormatError("XML: deformation entry must have 4 weights");
    }
  const XNode* idx = root->child("indexers");
  if (!idx) throw FormatError("XML: missing <indexers>");
  for (size_t c = 0; c < idx->kids.size(); ++c) {
    const XNode* cn = idx->child(("component-" + std::to_string(c)).c_str());
    if (!cn) throw FormatError("XML: missing component-" + std::to_string(c));
    std::vector<Part> parts;
    for (size_t p = 0; p < cn->kids.size(); ++p) {
      const XNode* pn = cn->child(("part-" + std::to_string(p)).c_str());
      if (!pn) throw FormatError("XML: missing part-" + std::to_string(p));
      Part P;
      P.parentid = (int)scalar(pn, "parentid");
      for_each_number(pn->child("filterid"), [&](double x) { P.filterid.push_back((int)x); });
      for_each_number(pn->child("biasid"), [&](double x) { P.biasid.push_back((int)x); });
      for_each_number(pn->child("defid"), [&](double x) { P.defid.push_back((int)x); });
      if (P.defid.empty()) P.defid.push_back(0);      // root: <defid></defid> (reference pushes 0, :151)
      parts.push_back(std::move(P));
    }
    m.comps.push_back(std::move(parts));
  }
  m.validate();
}

// ---- opencv_storage YAML (cv::FileStorage picks the format by file extension; %YAML:1.0 subset it writes) ----------
namespace {
struct YLine { int indent; char* b; char* e; };      // content without indentation / trailing white space

struct YParser {
  std::vector<YLine> lines;
  size_t i = 0;
  [[noreturn]] static void fail(const std::string& what) { throw FormatError("YAML: " + what); }

  explicit YParser(std::string& text) {
    char* p = &text[0];
    char* end = p + text.size();
    while (p < end) {
      char* nl = (char*)memchr(p, '\n', end - p);
      char* le = nl ? nl : end;
      char* b = p;
      while (b < le && *b == ' ') ++b;
      char* e = le;
      while (e > b && (e[-1] == ' ' || e[-1] == '\r' || e[-1] == '\t')) --e;
      const bool skip = b == e || *b == '#' || *b == '%' || (e - b == 3 && (!strncmp(b, "---", 3) || !strncmp(b, "...", 3)));
      if (!skip) lines.push_back({(int)(b - p), b, e});
      p = nl ? nl + 1 : end;
    }
  }
  // value text starting at v on the current line (index li): a flow collection may continue over the following lines; brackets
  // and commas are blanked so that the numeric scanner sees white-space separated values
  void leaf(XNode* n, char* v, char* le) {
    if (v < le && (*v == '[' || *v == '{')) {
      int depth = 0;
      char* tb = v + 1;
      for (;;) {
        for (char* c = v; c < le; ++c) {
          if (*c == '[' || *c == '{') { ++depth; *c = ' '; }
          else if (*c == ']' || *c == '}') { --depth; *c = ' '; if (depth == 0) { n->tb = tb; n->te = c; return; } }
          else if (*c == ',') *c = ' ';
        }
        *le = ' ';                                    // join the continuation line (the line break is ordinary white space)
        if (++i >= lines.size()) fail("unterminated flow sequence");
        v = lines[i].b; le = lines[i].e;
      }
    }
    n->tb = v; n->te = le;
  }
  // after "key:" / "- " the rest of the line: optional !!tag, then a scalar / flow value, or nothing (a nested block follows)
  void value(XNode* n, char* v, char* le, int indent) {
    while (v < le && *v == ' ') ++v;
    if (le - v >= 2 && v[0] == '!' && v[1] == '!') {
      char* t = v + 2;
      while (v < le && *v != ' ') ++v;
      n->type_id.assign(t, v);
      while (v < le && *v == ' ') ++v;
    }
    if (v < le) { leaf(n, v, le); ++i; return; }
    ++i;
    if (i < lines.size() && (lines[i].indent > indent || (lines[i].indent == indent && lines[i].e - lines[i].b >= 1 && lines[i].b[0] == '-' &&
                                                          (lines[i].e - lines[i].b == 1 || lines[i].b[1] == ' '))))
      block(n, lines[i].indent);
    else { n->tb = n->te = le; }                       // empty value
  }
  static bool is_dash(const YLine& l) { return *l.b == '-' && (l.e - l.b == 1 || l.b[1] == ' '); }
  void block(XNode* parent, int indent) {
    const bool seq = i < lines.size() && is_dash(lines[i]);
    while (i < lines.size() && lines[i].indent >= indent) {
      if (lines[i].indent > indent) fail("unexpected indentation");
      if (is_dash(lines[i]) != seq) {                   // a sequence written at its key's own indentation ends at the next key
        if (seq) return;
        fail("sequence item inside a mapping");
      }
      char* b = lines[i].b;
      char* e = lines[i].e;
      auto n = std::make_unique<XNode>();
      if (seq) {                                        // sequence item
        n->name = "_";
        XNode* raw = n.get();
        parent->kids.push_back(std::move(n));
        // the item's nested map (if any) is indented relative to the dash
        value(raw, b + 1, e, indent);
      } else {
        char* c = b;
        while (c < e && !(*c == ':' && (c + 1 == e || c[1] == ' '))) ++c;
        if (c == e) fail("expected 'key: value' in line '" + std::string(b, e) + "'");
        n->name.assign(b, c);
        XNode* raw = n.get();
        parent->kids.push_back(std::move(n));
        value(raw, c + 1, e, indent);
      }
    }
  }
};

void yaml_number(FILE* f, double v) {
  if (std::isnan(v)) { fputs(".Nan", f); return; }
  if (std::isinf(v)) { fputs(v > 0 ? ".Inf" : "-.Inf", f); return; }
  char buf[40];
  snprintf(buf, sizeof(buf), "%.17g", v);
  fputs(buf, f);
  if (!strpbrk(buf, ".eE")) fputc('.', f);             // cv::FileStorage marks reals with a decimal point ("3.")
}
}  // namespace

bool is_yaml_path(const std::string& path) {
  auto ends = [&](const char* suf) { const size_t n = strlen(suf); return path.size() >= n && !strcasecmp(path.c_str() + path.size() - n, suf); };
  return ends(".yml") || ends(".yaml");
}

void load_yaml(const std::string& path, Model& m) {
  std::string text = slurp(path);
  text.push_back('\n');
  YParser yp(text);
  XNode root;
  root.name = "opencv_storage";
  if (!yp.lines.empty()) yp.block(&root, yp.lines[0].indent);
  model_from_tree(&root, m);
}

void load_storage(const std::string& path, Model& m) {
  const std::string text = slurp(path);
  size_t a = text.find_first_not_of(" \n\r\t");
  if (a != std::string::npos && text[a] == '<') load_xml(path, m); else load_yaml(path, m);
}

void save_yaml(const Model& m, const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw IoError("cannot open '" + path + "' for writing: " + strerror(errno));
  bool plain = !m.name.empty();
  for (char ch : m.name) if (!isalnum((unsigned char)ch) && ch != '_' && ch != '-') plain = false;
  fprintf(f, "%%YAML:1.0\n---\n");
  if (plain) fprintf(f, "name: %s\n", m.name.c_str()); else fprintf(f, "name: \"%s\"\n", m.name.c_str());
  fprintf(f, "interval: %d\nthresh: ", m.interval);
  yaml_number(f, (double)m.thresh);
  fprintf(f, "\nsbin: %d\nnorient: %d\nflen: %d\nfiltersw:\n", m.sbin, m.norient, m.flen);
  for (int i = 0; i < m.nfilters(); ++i) {
    fprintf(f, "   - !!opencv-matrix\n      rows: %d\n      cols: %d\n      dt: d\n      data: [ ", m.frows[i], m.fkw[i] * m.flen);
    for (size_t j = 0; j < m.filters[i].size(); ++j) {
      if (j) fputs(j % 3 ? ", " : ",\n          ", f);
      yaml_number(f, m.filters[i][j]);
    }
    fputs(" ]\n", f);
  }
  auto flow_f = [&](const char* key, const float* v, size_t n, const char* cont) {
    fprintf(f, "%s[ ", key);
    for (size_t j = 0; j < n; ++j) { if (j) fputs(j % 4 ? ", " : cont, f); yaml_number(f, (double)v[j]); }
    fputs(n ? " ]\n" : "]\n", f);
  };
  auto flow_i = [&](const char* key, const int* v, size_t n, const char* cont) {
    fprintf(f, "%s[", key);
    for (size_t j = 0; j < n; ++j) fprintf(f, "%s%d", j ? (j % 16 ? ", " : cont) : " ", v[j]);
    fputs(n ? " ]\n" : "]\n", f);
  };
  flow_f("biasw: ", m.biasw.data(), m.biasw.size(), ",\n    ");
  flow_i("anchors: ", m.anchors.data(), m.anchors.size(), ",\n    ");
  fputs("defs:\n", f);
  for (int j = 0; j < m.ndefs(); ++j) flow_f("   - ", m.defs.data() + 4 * j, 4, ",\n       ");
  fputs("indexers:\n", f);
  for (size_t c = 0; c < m.comps.size(); ++c) {
    fprintf(f, "   component-%zu:\n", c);
    for (size_t p = 0; p < m.comps[c].size(); ++p) {
      const Part& P = m.comps[c][p];
      fprintf(f, "      part-%zu:\n         parentid: %d\n", p, P.parentid);
      flow_i("         filterid: ", P.filterid.data(), P.filterid.size(), ",\n             ");
      flow_i("         biasid: ", P.biasid.data(), P.biasid.size(), ",\n             ");
      if (p == 0) fputs("         defid: []\n", f); else flow_i("         defid: ", P.defid.data(), P.defid.size(), ",\n             ");
    }
  }
  if (fclose(f) != 0) throw IoError("write failed on '" + path + "'");
}

void save_storage(const Model& m, const std::string& path) {
  if (is_yaml_path(path)) save_yaml(m, path); else save_xml(m, path);
}

void save_xml(const Model& m, const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw IoError("cannot open '" + path + "' for writing: " + strerror(errno));
  fprintf(f, "<?xml version=\"1.0\"?>\n<opencv_storage>\n");
  bool plain = !m.name.empty();
  for (char ch : m.name) if (!isalnum((unsigned char)ch) && ch != '_' && ch != '-') plain = false;
  std::string nm;
  for (char ch : m.name) { if (ch == '&') nm += "&amp;"; else if (ch == '<') nm += "&lt;"; else if (ch == '>') nm += "&gt;"; else nm += ch; }
  if (!plain) nm = "\"" + nm + "\"";       // cv::FileStorage quotes strings that are not plain identifiers
  fprintf(f, "<name>%s</name>\n<interval>%d</interval>\n<thresh>%.8e</thresh>\n<sbin>%d</sbin>\n<norient>%d</norient>\n<flen>%d</flen>\n",
          nm.c_str(), m.interval, (double)m.thresh, m.sbin, m.norient, m.flen);
  fprintf(f, "<filtersw>\n");
  for (int i = 0; i < m.nfilters(); ++i) {
    fprintf(f, "  <_ type_id=\"opencv-matrix\">\n    <rows>%d</rows>\n    <cols>%d</cols>\n    <dt>d</dt>\n    <data>\n", m.frows[i], m.fkw[i] * m.flen);
    for (size_t j = 0; j < m.filters[i].size(); ++j) fprintf(f, "%s%.16e", j % 2 ? " " : "\n      ", m.filters[i][j]);
    fprintf(f, "</data></_>\n");
  }
  fprintf(f, "</filtersw>\n<biasw>");
  for (size_t j = 0; j < m.biasw.size(); ++j) fprintf(f, "%s%.8e", j % 4 ? " " : "\n  ", (double)m.biasw[j]);
  fprintf(f, "</biasw>\n<anchors>");
  for (size_t j = 0; j < m.anchors.size(); ++j) fprintf(f, "%s%d", j % 24 ? " " : "\n  ", m.anchors[j]);
  fprintf(f, "</anchors>\n<defs>\n");
  for (int j = 0; j < m.ndefs(); ++j)
    fprintf(f, "  <_>\n    %.8e %.8e %.8e %.8e</_>\n", (double)m.defs[j * 4], (double)m.defs[j * 4 + 1], (double)m.defs[j * 4 + 2], (double)m.defs[j * 4 + 3]);
  fprintf(f, "</defs>\n<indexers>\n");
  auto seq = [&](const char* tag, const std::vector<int>& v, bool empty_ok) {
    if (v.empty() && empty_ok) { fprintf(f, "      <%s></%s>", tag, tag); return; }
    if (v.size() == 1) { fprintf(f, "      <%s>%d</%s>\n", tag, v[0], tag); return; }
    fprintf(f, "      <%s>\n       ", tag);
    for (int x : v) fprintf(f, " %d", x);
    fprintf(f, "</%s>\n", tag);
  };
  for (size_t c = 0; c < m.comps.size(); ++c) {
    fprintf(f, "  <component-%zu>\n", c);
    for (size_t p = 0; p < m.comps[c].size(); ++p) {
      const Part& P = m.comps[c][p];
      fprintf(f, "    <part-%zu>\n      <parentid>%d</parentid>\n", p, P.parentid);
      seq("filterid", P.filterid, false);
      seq("biasid", P.biasid, false);
      if (p == 0) fprintf(f, "      <defid></defid>"); else seq("defid", P.defid, false);
      fprintf(f, "</part-%zu>\n", p);
    }
    fprintf(f, "  </component-%zu>\n", c);
  }
  fprintf(f, "</indexers>\n</opencv_storage>\n");
  if (fclose(f) != 0) throw IoError("write failed on '" + path + "'");
}

// ---- binary container -------------------------------------------------------
namespace {
struct Writer {
  FILE* f;
  template <typename T> void put(const T& v) { fwrite(&v, sizeof(T), 1, f); }
  template <typename T> void vec(const std::vector<T>& v) { put<uint32_t>((uint32_t)v.size()); if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f); }
};
struct Reader {
  const std::string& s;
  size_t o = 0;
  template <typename T> T get() { if (o + sizeof(T) > s.size()) throw FormatError("PBDM: truncated"); T v; memcpy(&v, s.data() + o, sizeof(T)); o += sizeof(T); return v; }
  template <typename T> void vec(std::vector<T>& v) {
    const uint32_t n = get<uint32_t>();
    if (o + (size_t)n * sizeof(T) > s.size()) throw FormatError("PBDM: truncated");
    v.resize(n);
    if (n) memcpy(v.data(), s.data() + o, (size_t)n * sizeof(T));
    o += (size_t)n * sizeof(T);
  }
};
}  // namespace

void save_bin(const Model& m, const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw IoError("cannot open '" + path + "' for writing: " + strerror(errno));
  Writer w{f};
  fwrite("PBDM", 1, 4, f);
  w.put<uint32_t>(1);
  w.put<uint32_t>((uint32_t)m.name.size());
  fwrite(m.name.data(), 1, m.name.size(), f);
  w.put<int32_t>(m.interval); w.put<float>(m.thresh); w.put<int32_t>(m.sbin); w.put<int32_t>(m.norient); w.put<int32_t>(m.flen);
  w.put<uint32_t>((uint32_t)m.nfilters());
  for (int i = 0; i < m.nfilters(); ++i) { w.put<int32_t>(m.frows[i]); w.put<int32_t>(m.fkw[i]); w.vec(m.filters[i]); }
  w.vec(m.biasw); w.vec(m.anchors); w.vec(m.defs);
  w.put<uint32_t>((uint32_t)m.comps.size());
  for (auto& c : m.comps) {
    w.put<uint32_t>((uint32_t)c.size());
    for (auto& P : c) { w.put<int32_t>(P.parentid); w.vec(P.filterid); w.vec(P.biasid); w.vec(P.defid); }
  }
  if (fclose(f) != 0) throw IoError("write failed on '" + path + "'");
}

void load_bin(const std::string& path, Model& m) {
  const std::string s = slurp(path);
  if (s.size() < 8 || memcmp(s.data(), "PBDM", 4)) throw FormatError("not a PBDM file: '" + path + "'");
  Reader r{s, 4};
  if (r.get<uint32_t>() != 1) throw FormatError("PBDM: unsupported version");
  m = Model();
  const uint32_t nl = r.get<uint32_t>();
  if (r.o + nl > s.size()) throw FormatError("PBDM: truncated");
  m.name.assign(s.data() + r.o, nl); r.o += nl;
  m.interval = r.get<int32_t>(); m.thresh = r.get<float>(); m.sbin = r.get<int32_t>(); m.norient = r.get<int32_t>(); m.flen = r.get<int32_t>();
  // every count is bounded by the bytes that are left (each filter / component / part occupies at least 4 bytes)
  const uint32_t nf = r.get<uint32_t>();
  if (nf > (s.size() - r.o) / 4) throw FormatError("PBDM: filter count exceeds the file size");
  for (uint32_t i = 0; i < nf; ++i) {
    m.frows.push_back(r.get<int32_t>()); m.fkw.push_back(r.get<int32_t>());
    std::vector<double> v; r.vec(v); m.filters.push_back(std::move(v));
  }
  r.vec(m.biasw); r.vec(m.anchors); r.vec(m.defs);
  const uint32_t nc = r.get<uint32_t>();
  if (nc > (s.size() - r.o) / 4) throw FormatError("PBDM: component count exceeds the file size");
  for (uint32_t c = 0; c < nc; ++c) {
    const uint32_t np = r.get<uint32_t>();
    if (np > (s.size() - r.o) / 4) throw FormatError("PBDM: part count exceeds the file size");
    std::vector<Part> parts(np);
    for (auto& P : parts) { P.parentid = r.get<int32_t>(); r.vec(P.filterid); r.vec(P.biasid); r.vec(P.defid); }
    m.comps.push_back(std::move(parts));
  }
  m.validate();
}

}  // namespace pbd
