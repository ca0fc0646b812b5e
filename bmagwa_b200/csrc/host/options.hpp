// options.hpp -- configuration: the reference's INI file, key for key.
//
// Mirrors bmagwa::Options (src/options.hpp:138-321,324-459) over an INI reader with the semantics
// of the inih parser the reference vendors (src/inih/ini.c:60-140, cpp/INIReader.cpp:36-49):
// [section] headers, name=value or name:value, ';' / '#' comment lines, inline " ;" comments,
// case-insensitive "section.name" lookup, unknown keys ignored, lines longer than 199 characters
// cut.  Same defaults, same validation, same error messages.  Extra keys understood by this
// implementation live in section [b200] (tau_rng, device) so reference INI files stay valid.
#pragma once
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace bmg {

class IniFile {
 public:
  explicit IniFile(const std::string& path)
  {
    std::ifstream in(path.c_str());
    if (!in.good()) { error_ = -1; return; }
    std::string raw, section, last_name;
    int lineno = 0;
    while (std::getline(in, raw)) {
      ++lineno;
      if (raw.size() > 199) raw.resize(199);  // MAX_LINE 200 incl. terminator
      std::string line = rstrip(raw);
      const size_t first = line.find_first_not_of(" \t\r\n\v\f");
      const std::string body = first == std::string::npos ? std::string() : line.substr(first);
      if (!last_name.empty() && !body.empty() && first > 0) {
        values_[key(section, last_name)] = body;  // continuation line replaces the value (ini.c:88-93)
        continue;
      }
      if (body.empty() || body[0] == ';' || body[0] == '#') continue;
      if (body[0] == '[') {
        const size_t close = find_or_comment(body, 1, ']');
        if (close < body.size() && body[close] == ']') { section = body.substr(1, close - 1); last_name.clear(); }
        else if (!error_) error_ = lineno;
        continue;
      }
      size_t sep = find_or_comment(body, 0, '=');
      if (sep >= body.size() || body[sep] != '=') sep = find_or_comment(body, 0, ':');
      if (sep < body.size() && (body[sep] == '=' || body[sep] == ':')) {
        const std::string name = rstrip(body.substr(0, sep));
        std::string value = body.substr(sep + 1);
        const size_t vs = value.find_first_not_of(" \t\r\n\v\f");
        value = vs == std::string::npos ? std::string() : value.substr(vs);
        const size_t cm = find_or_comment(value, 0, '\0');
        if (cm < value.size() && value[cm] == ';') value.resize(cm);
        value = rstrip(value);
        last_name = name;
        values_[key(section, name)] = value;
      } else if (!error_) error_ = lineno;
    }
  }
  int parse_error() const { return error_; }
  std::string get(const std::string& section, const std::string& name, const std::string& dflt) const
  {
    auto it = values_.find(key(section, name));
    return it == values_.end() ? dflt : it->second;
  }
  void set(const std::string& section, const std::string& name, const std::string& value) { values_[key(section, name)] = value; }

 private:
  static std::string rstrip(std::string s)
  {
    while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
    return s;
  }
  // position of `c`, or of a ';' that follows whitespace, whichever comes first (ini.c:37-47)
  static size_t find_or_comment(const std::string& s, size_t from, char c)
  {
    bool was_space = false;
    size_t i = from;
    for (; i < s.size(); ++i) {
      if (c != '\0' && s[i] == c) break;
      if (was_space && s[i] == ';') break;
      was_space = std::isspace((unsigned char)s[i]) != 0;
    }
    return i;
  }
  static std::string key(const std::string& section, const std::string& name)
  {
    std::string k = section + "." + name;
    for (char& ch : k) ch = (char)std::tolower((unsigned char)ch);
    return k;
  }
  std::map<std::string, std::string> values_;
  int error_ = 0;
};

enum EffectType { kA = 0, kH = 1, kD = 2, kR = 3, kAH = 4 };

struct Options {
  std::string file_fam, file_g, file_e, file_y;
  size_t n = 0, m_g = 0, m_e = 0;
  bool recode_g_to_minor_allele_count = false;
  std::vector<int> types;
  int sampler_type = 0;
  size_t do_n_iter = 0, n_rao = 0, n_rao_burnin = 0, thin = 0, n_sample_tau2_and_missing = 0;
  size_t max_move_size = 20, max_SNP_neighborhood_size = 10;
  double p_move_size = 0.2, p_move_size_nbs = 0.2, p_move_size_nbc = 0.2;
  bool adapt_p_move_size = true;
  double p_move_size_acpt_goal = 0.0;
  bool flat_proposal_dist = false, adaptation = false, save_beta = false;
  size_t verbosity = 0;
  std::string basename = "chain";
  size_t n_threads = 1;
  std::vector<uint32_t> seeds;
  double e_qg = 0, var_qg = 0, nu_sigma2 = 0, s2_sigma2 = 0, R2mode_sigma2 = 0;
  double nu_tau2[4] = {0, 0, 0, 0}, s2_tau2[4] = {0, 0, 0, 0}, eh_tau2[4] = {0, 0, 0, 0}, mu_alpha = 0;
  double inv_tau2_e_const_val = 0, inv_tau2_e_val = 0;
  double types_prior[5] = {1, 1, 1, 1, 1};
  bool use_individual_tau2 = false;
  size_t delay_rejection = 0;
  // [b200] extensions
  bool probit = false;            // y holds 0/1 case-control labels; the sampler works on an Albert-Chib latent phenotype
  std::string tau_rng = "host";   // "host": per-SNP tau draws from the chain's stream in reference order (parity);
                                  // "device": counter-based draws on the GPU (throughput)
  int device = 0;
  size_t pip_burnin = 0;          // thinned samples dropped before the running inclusion counts start ([b200] pip_burnin)
  bool quiet = false;

  explicit Options(const std::string& path, bool quiet_ = false) : quiet(quiet_)
  {
    IniFile r(path);
    if (r.parse_error() != 0) throw std::runtime_error("Cannot load/parse configuration file.");
    parse(r);
  }

 private:
  void warn(const std::string& dflt, const std::string& s, const std::string& n) const
  {
    if (!quiet) std::cout << "Warning: using default value of " << dflt << " for option " << s << "." << n << std::endl;
  }
  template <class T>
  static std::string show(T v) { std::ostringstream o; o << v; return o.str(); }
  template <class T>
  static T convert(const std::string& s)
  {
    std::istringstream i(s);
    T x;
    char c;
    if (!(i >> x) || i.get(c)) throw std::runtime_error("Invalid conversion.");
    return x;
  }
  std::string str(IniFile& r, const char* s, const char* n, bool allow, const std::string& d) const
  {
    const std::string v = r.get(s, n, "DEFAULT_VALUE");
    const bool none = v.empty() || v == "DEFAULT_VALUE";
    if (none && !allow) throw std::runtime_error(std::string("Cannot leave option empty: ") + s + "." + n);
    if (none) { warn(d, s, n); return d; }
    return v;
  }
  double dbl(IniFile& r, const char* s, const char* n, bool allow, double d) const
  {
    const std::string v = r.get(s, n, "DEFAULT_VALUE");
    const bool none = v.empty() || v == "DEFAULT_VALUE";
    if (none && !allow) throw std::runtime_error(std::string("Cannot leave option empty: ") + s + "." + n);
    if (none) { warn(show(d), s, n); return d; }
    return convert<double>(v);
  }
  double dbl01(IniFile& r, const char* s, const char* n, bool allow, double d) const
  {
    const double v = dbl(r, s, n, allow, d);
    if (v < 0 || v > 1) throw std::runtime_error("Value in 0..1 required");
    return v;
  }
  double dblpos(IniFile& r, const char* s, const char* n, bool allow, double d) const
  {
    const double v = dbl(r, s, n, allow, d);
    if (v <= 0) throw std::runtime_error(std::string("Positive value required for ") + s + "." + n);
    return v;
  }
  double dblnn(IniFile& r, const char* s, const char* n, bool allow, double d) const
  {
    const double v = dbl(r, s, n, allow, d);
    if (v < 0) throw std::runtime_error(std::string("Non-negative value required for ") + s + "." + n);
    return v;
  }
  int integer(IniFile& r, const char* s, const char* n, bool allow, int d) const
  {
    const std::string v = r.get(s, n, "DEFAULT_VALUE");
    const bool none = v.empty() || v == "DEFAULT_VALUE";
    if (none && !allow) throw std::runtime_error(std::string("Cannot leave option empty: ") + s + "." + n);
    if (none) { warn(show(d), s, n); return d; }
    return convert<int>(v);
  }
  int intnn(IniFile& r, const char* s, const char* n, bool allow, int d) const
  {
    const int v = integer(r, s, n, allow, d);
    if (v < 0) throw std::runtime_error("Nonnegative integer required");
    return v;
  }
  int intpos(IniFile& r, const char* s, const char* n, bool allow, int d) const
  {
    const int v = integer(r, s, n, allow, d);
    if (v <= 0) throw std::runtime_error("Positive integer required");
    return v;
  }
  bool boolean(IniFile& r, const char* s, const char* n, bool allow, int d) const
  {
    const int v = integer(r, s, n, allow, d);
    if (v != 0 && v != 1) throw std::runtime_error("Boolean value required (0 or 1)");
    return v == 1;
  }
  static std::vector<std::string> split_commas(const std::string& s)
  {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) {
      if (ch == ',') { out.push_back(cur); cur.clear(); }
      else cur += ch;
    }
    out.push_back(cur);
    for (std::string& t : out) {
      size_t a = 0, b = t.size();
      while (a < b && std::isspace((unsigned char)t[a])) ++a;
      while (b > a && std::isspace((unsigned char)t[b - 1])) --b;
      t = t.substr(a, b - a);
    }
    return out;
  }
  bool has_type(int t) const { return std::find(types.begin(), types.end(), t) != types.end(); }
  void tau_prior(IniFile& r, int t, const char* nu, const char* eh, const char* s2, const char* msg)
  {
    nu_tau2[t] = dblpos(r, "prior", nu, false, 0.0);
    eh_tau2[t] = dbl01(r, "prior", eh, true, 0.0);
    s2_tau2[t] = dblnn(r, "prior", s2, true, 0.0);
    if (s2_tau2[t] <= 0 && eh_tau2[t] <= 0) throw std::runtime_error(msg);
  }

  void parse(IniFile& r)
  {
    // same order as src/options.hpp:140-321 so that warnings and the first error raised agree
    n = intpos(r, "sizes", "n", false, 0);
    m_g = intpos(r, "sizes", "m_g", false, 0);
    m_e = intnn(r, "sizes", "m_e", true, 0);
    file_fam = str(r, "datafiles", "file_fam", false, "");
    file_g = str(r, "datafiles", "file_g", false, "");
    file_e = str(r, "datafiles", "file_e", m_e == 0, "");
    file_y = str(r, "datafiles", "file_y", true, "");
    recode_g_to_minor_allele_count = boolean(r, "datafiles", "recode_g_to_minor_allele_count", true, 0);
    const std::string types_str = str(r, "model", "types", false, "");
    const std::string stype = str(r, "sampler", "type", true, "PMV");
    do_n_iter = intpos(r, "sampler", "do_n_iter", false, 0);
    n_rao = intnn(r, "sampler", "n_rao", false, 0);
    n_rao_burnin = intnn(r, "sampler", "n_rao_burnin", false, 0);
    thin = intpos(r, "sampler", "thin", false, 0);
    n_sample_tau2_and_missing = intpos(r, "sampler", "n_sample_tau2_and_missing", false, 0);
    max_move_size = intpos(r, "sampler", "max_move_size", true, 20);
    max_SNP_neighborhood_size = intpos(r, "sampler", "max_SNP_neighborhood_size", true, 10);
    p_move_size = dblpos(r, "sampler", "p_move_size", true, 0.2);
    if (p_move_size > 1) throw std::runtime_error("Config error: p_move_size must be <= 1.");
    p_move_size_nbs = dblpos(r, "sampler", "p_move_size_nbs", true, 0.2);
    if (p_move_size_nbs > 1) throw std::runtime_error("Config error: p_move_size_nbs must be <= 1.");
    p_move_size_nbc = dblpos(r, "sampler", "p_move_size_nbc", true, 0.2);
    if (p_move_size_nbc > 1) throw std::runtime_error("Config error: p_move_size_nbc must be <= 1.");
    adapt_p_move_size = boolean(r, "sampler", "adapt_p_move_size", true, 1);
    p_move_size_acpt_goal = dblnn(r, "sampler", "p_move_size_acpt_goal", true, 0.0);
    if (p_move_size_acpt_goal > 1) throw std::runtime_error("Config error: p_move_size_acpt_goal must be <= 1.");
    adaptation = boolean(r, "sampler", "adaptation", true, 0);
    flat_proposal_dist = boolean(r, "sampler", "flat_proposal_dist", true, 0);
    save_beta = boolean(r, "sampler", "save_beta", true, 0);
    verbosity = intnn(r, "sampler", "verbosity", false, 0);
    delay_rejection = intnn(r, "sampler", "delay_rejection", true, 0);
    basename = str(r, "thread", "basename", true, "chain");
    n_threads = intpos(r, "thread", "n_threads", true, 1);
    const std::string seeds_str = str(r, "thread", "seeds", false, "");
    e_qg = dblpos(r, "prior", "e_qg", false, 0.0);
    var_qg = dblpos(r, "prior", "var_qg", false, 0.0);
    nu_sigma2 = dblpos(r, "prior", "nu_sigma2", false, 0.0);
    R2mode_sigma2 = dbl01(r, "prior", "R2mode_sigma2", true, 0.0);
    s2_sigma2 = dblnn(r, "prior", "s2_sigma2", true, 0.0);
    if (s2_sigma2 <= 0 && R2mode_sigma2 <= 0) throw std::runtime_error("Config error: s2_sigma2 or R2mode_sigma2 must be > 0.");
    mu_alpha = dbl(r, "prior", "mu_alpha", true, 0.0);
    inv_tau2_e_const_val = dblnn(r, "prior", "inv_tau2_e_const_val", false, 0.0);
    inv_tau2_e_val = dblnn(r, "prior", "inv_tau2_e_val", m_e == 0, 0.0);
    use_individual_tau2 = boolean(r, "prior", "use_individual_tau2", true, 0);
    types_prior[kA] = dbl(r, "prior", "type_A", true, 1.0);
    types_prior[kH] = dbl(r, "prior", "type_H", true, 1.0);
    types_prior[kR] = dbl(r, "prior", "type_R", true, 1.0);
    types_prior[kD] = dbl(r, "prior", "type_D", true, 1.0);
    types_prior[kAH] = dbl(r, "prior", "type_AH", true, 1.0);

    if (stype == "PMV") sampler_type = 0;
    else if (stype == "NK") sampler_type = 1;
    else if (stype == "KSC") sampler_type = 2;
    else if (stype == "G") sampler_type = 3;
    else throw std::runtime_error("Config error: unknown sampler type");
    if (n_rao % thin != 0) throw std::runtime_error("Config error: n_rao should be a multiple of thin");

    for (const std::string& t : split_commas(types_str)) {
      if (t == "A") types.push_back(kA);
      else if (t == "H") types.push_back(kH);
      else if (t == "D") types.push_back(kD);
      else if (t == "R") types.push_back(kR);
      else if (t == "AH") types.push_back(kAH);
      else throw std::runtime_error("Config error: Unknown model type");
    }
    std::sort(types.begin(), types.end());
    if (types.empty()) throw std::runtime_error("Config error: Types cannot be empty");
    if (sampler_type != 0 && (types.size() != 1 || types[0] != kA))
      throw std::runtime_error("Config error: NK/KSC/G samplers support only A effect types");
    if (has_type(kA) || has_type(kAH)) tau_prior(r, kA, "nu_tau2_A", "eh_tau2_A", "s2_tau2_A", "Config error: s2_tau2_A or eh_tau2_A must be > 0.");
    if (has_type(kH) || has_type(kAH)) tau_prior(r, kH, "nu_tau2_H", "eh_tau2_H", "s2_tau2_H", "Config error: s2_tau2_H or eh_tau2_H must be > 0.");
    if (has_type(kD)) tau_prior(r, kD, "nu_tau2_D", "eh_tau2_D", "s2_tau2_D", "Config error: s2_tau2_D or eh_tau2_D must be > 0.");
    if (has_type(kR)) tau_prior(r, kR, "nu_tau2_R", "eh_tau2_R", "s2_tau2_R", "Config error: s2_tau2_R or eh_tau2_R must be > 0.");

    for (const std::string& t : split_commas(seeds_str)) seeds.push_back(convert<uint32_t>(t));
    if (seeds.size() != n_threads) throw std::runtime_error("Config error: Number of seeds must equal n_threads");
    for (size_t i = 0; i + 1 < n_threads; ++i)
      for (size_t j = i + 1; j < n_threads; ++j)
        if (seeds[i] == seeds[j]) throw std::runtime_error("Seeds are not unique");

    // [b200] extensions (never required)
    const std::string tr = r.get("b200", "tau_rng", "host");
    if (tr != "host" && tr != "device") throw std::runtime_error("Config error: b200.tau_rng must be host or device");
    tau_rng = tr;
    probit = r.get("b200", "probit", "0") != "0";
    pip_burnin = convert<size_t>(r.get("b200", "pip_burnin", "0"));
    const std::string dv = r.get("b200", "device", "0");
    device = convert<int>(dv);
  }
};

}  // namespace bmg
