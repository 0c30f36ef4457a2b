// host_smps.cuh -- SMPS (.cor / .tim / .sto) reader and the stage-2 tables the device path consumes
// (SURVEY.md 8(f) row N4).  Host-only C++; part of the single translation unit sqlp_api.cu.
//
// What it follows in the reference (behaviour, not code):
//   cor  src/smps/smps_cor.jl:26-194   sections NAME ROWS COLUMNS RHS BOUNDS ENDATA, '*' comments, a
//        line that starts in column 1 is a section header, later COLUMNS entries overwrite earlier
//        ones, bound types LO UP FX FR MI PL, defaults [0, +Inf), the first row must be 'N' (:178-179)
//   tim  src/smps/smps_tim.jl:30-64    implicit PERIODS lines "col row period"
//   sto  src/smps/smps_sto.jl:41-111   INDEP DISCRETE | NORMAL | UNIFORM, one keyword only
//   stage split  src/smps/smps_prob.jl:14-102   stage-2 rows / columns start at the second period's
//        row / column; the template keeps "func (>=, <=, ==) rhs" so Tbar is the cor block, rbar the rhs
//   coefficient extraction  src/sd_algorithm/subprob.jl:15-69   exact zeros are not stored
// The reference keeps the random elements in a Dict (hash order, smps_sto.jl:35); the element order here
// is fixed once: order of first appearance in the .sto file.
#pragma once

#include <cerrno>
#include <cmath>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <unordered_map>

namespace {

struct SmpsLine {
    bool header;
    std::vector<std::string> tok;
    int lineno;
};

inline std::vector<SmpsLine> smps_lines(const char *path, const char *what)
{
    REQUIRE(path && *path, SQLP_E_INVALID, std::string("null ") + what + " path");
    std::ifstream in(path);
    REQUIRE(in.good(), SQLP_E_IO, std::string("cannot open ") + what + " file " + path);
    std::vector<SmpsLine> out;
    std::string raw;
    int no = 0;
    while (std::getline(in, raw)) {
        ++no;
        if (!raw.empty() && raw.back() == '\r') raw.pop_back();
        if (raw.empty() || raw[0] == '*') continue;
        SmpsLine l;
        l.header = !(raw[0] == ' ' || raw[0] == '\t');
        l.lineno = no;
        std::istringstream ss(raw);
        std::string t;
        while (ss >> t) l.tok.push_back(t);
        if (l.tok.empty()) continue;          // blanks only
        out.push_back(std::move(l));
    }
    return out;
}

inline std::string smps_where(const char *path, const SmpsLine &l)
{
    return std::string(path) + ":" + std::to_string(l.lineno);
}

inline double smps_number(const std::string &t, const char *path, const SmpsLine &l)
{
    errno = 0;
    char *end = nullptr;
    double v = strtod(t.c_str(), &end);
    REQUIRE(end != t.c_str() && *end == 0, SQLP_E_INVALID, smps_where(path, l) + ": '" + t + "' is not a number");
    return v;
}

struct SmpsElement {
    std::string col, row;
    int kind;                      // 0 DISCRETE, 1 NORMAL(mean, variance), 2 UNIFORM(left, right)
    double a = 0.0, b = 0.0;
    std::vector<double> val, prob;
};

}  // namespace

struct sqlp_smps {
    // cor
    std::string name;
    std::vector<char> dir;
    std::vector<std::string> rows, cols;
    std::unordered_map<std::string, int64_t> row_of, col_of;
    std::vector<std::map<int64_t, double>> col_entries;   // per column: row -> value (as written, zeros kept)
    std::vector<double> rhs, lower, upper;
    // tim
    std::string tim_name;
    std::vector<std::array<std::string, 3>> periods;       // (period, col, row)
    // sto
    std::string sto_name;
    bool has_sto = false;
    std::vector<SmpsElement> elems;
    // stage 2
    int64_t c2 = 0, r2 = 0, n1 = 0, n2 = 0, m2 = 0;
    std::vector<int64_t> T_colptr, T_rowval, W_colptr, W_rowval, r_idx;
    std::vector<double> T_nzval, W_nzval, r_val, cost, x_cost;
    std::vector<int32_t> pos_row, pos_col;
    int64_t max_outcomes = 0;
};

namespace {

inline void smps_read_cor(sqlp_smps &p, const char *path)
{
    std::string section;
    bool named = false;
    std::vector<const SmpsLine *> colsec, rhssec, bndsec;
    auto lines = smps_lines(path, "cor");
    for (const SmpsLine &l : lines) {
        if (l.header) {
            section = l.tok[0];
            REQUIRE(section == "NAME" || section == "ROWS" || section == "COLUMNS" || section == "RHS" ||
                        section == "BOUNDS" || section == "ENDATA",
                    SQLP_E_INVALID, smps_where(path, l) + ": unsupported cor section " + section);
            if (section == "NAME") {
                REQUIRE(l.tok.size() >= 2, SQLP_E_INVALID, smps_where(path, l) + ": NAME without a name");
                p.name = l.tok[1];
                named = true;
            }
        } else if (section == "ROWS") {
            REQUIRE(l.tok.size() >= 2, SQLP_E_INVALID, smps_where(path, l) + ": ROWS line needs 'direction name'");
            p.dir.push_back(l.tok[0][0]);
            p.row_of[l.tok[1]] = (int64_t)p.rows.size();      // a repeated name maps to its last row
            p.rows.push_back(l.tok[1]);
        } else if (section == "COLUMNS") {
            colsec.push_back(&l);
        } else if (section == "RHS") {
            rhssec.push_back(&l);
        } else if (section == "BOUNDS") {
            bndsec.push_back(&l);
        } else {
            throw Error(SQLP_E_INVALID, smps_where(path, l) + ": data line outside a section");
        }
    }
    REQUIRE(named, SQLP_E_INVALID, std::string(path) + ": no NAME section");
    REQUIRE(!p.rows.empty() && p.dir[0] == 'N', SQLP_E_INVALID,
            std::string(path) + ": first row of the cor file is not the objective (N) row");
    for (const SmpsLine *l : colsec)
        if (!p.col_of.count(l->tok[0])) {
            p.col_of[l->tok[0]] = (int64_t)p.cols.size();
            p.cols.push_back(l->tok[0]);
        }
    p.col_entries.resize(p.cols.size());
    auto row_index = [&](const std::string &r, const SmpsLine &l) {
        auto it = p.row_of.find(r);
        REQUIRE(it != p.row_of.end(), SQLP_E_INVALID, smps_where(path, l) + ": unknown row " + r);
        return it->second;
    };
    for (const SmpsLine *l : colsec) {
        REQUIRE(l->tok.size() % 2 == 1, SQLP_E_INVALID, smps_where(path, *l) + ": COLUMNS line needs (row, value) pairs");
        int64_t j = p.col_of[l->tok[0]];
        for (size_t t = 1; t + 1 < l->tok.size(); t += 2)
            p.col_entries[j][row_index(l->tok[t], *l)] = smps_number(l->tok[t + 1], path, *l);
    }
    p.rhs.assign(p.rows.size(), 0.0);
    for (const SmpsLine *l : rhssec) {
        REQUIRE(l->tok.size() % 2 == 1, SQLP_E_INVALID, smps_where(path, *l) + ": RHS line needs 'set (row value)...'");
        for (size_t t = 1; t + 1 < l->tok.size(); t += 2)
            p.rhs[row_index(l->tok[t], *l)] = smps_number(l->tok[t + 1], path, *l);
    }
    const double inf = std::numeric_limits<double>::infinity();
    p.lower.assign(p.cols.size(), 0.0);
    p.upper.assign(p.cols.size(), inf);
    for (const SmpsLine *l : bndsec) {
        REQUIRE(l->tok.size() >= 3, SQLP_E_INVALID, smps_where(path, *l) + ": BOUNDS line needs 'type set column [value]'");
        const std::string &bt = l->tok[0];
        auto it = p.col_of.find(l->tok[2]);
        REQUIRE(it != p.col_of.end(), SQLP_E_INVALID, smps_where(path, *l) + ": unknown column " + l->tok[2]);
        int64_t j = it->second;
        auto value = [&] {
            REQUIRE(l->tok.size() >= 4, SQLP_E_INVALID, smps_where(path, *l) + ": bound " + bt + " needs a value");
            return smps_number(l->tok[3], path, *l);
        };
        if (bt == "LO") p.lower[j] = value();
        else if (bt == "UP") p.upper[j] = value();
        else if (bt == "FX") p.lower[j] = p.upper[j] = value();
        else if (bt == "FR") { p.lower[j] = -inf; p.upper[j] = inf; }
        else if (bt == "MI") p.lower[j] = -inf;
        else if (bt == "PL") p.upper[j] = inf;
        else throw Error(SQLP_E_INVALID, smps_where(path, *l) + ": unsupported bound type " + bt + " for variable " + l->tok[2]);
    }
}

inline void smps_read_tim(sqlp_smps &p, const char *path)
{
    std::string section;
    for (const SmpsLine &l : smps_lines(path, "tim")) {
        if (l.header) {
            section = l.tok[0];
            REQUIRE(section == "TIME" || section == "PERIODS" || section == "ENDATA", SQLP_E_INVALID,
                    smps_where(path, l) + ": unsupported tim section " + section);
            if (section == "TIME" && l.tok.size() >= 2) p.tim_name = l.tok[1];
        } else {
            REQUIRE(section == "PERIODS", SQLP_E_INVALID, smps_where(path, l) + ": data line outside PERIODS");
            REQUIRE(l.tok.size() >= 3, SQLP_E_INVALID, smps_where(path, l) + ": PERIODS line needs 'column row period'");
            p.periods.push_back({l.tok[2], l.tok[0], l.tok[1]});
        }
    }
}

inline void smps_read_sto(sqlp_smps &p, const char *path)
{
    std::string section;
    std::vector<std::string> keys;
    std::map<std::pair<std::string, std::string>, size_t> at;
    for (const SmpsLine &l : smps_lines(path, "sto")) {
        if (l.header) {
            section = l.tok[0];
            REQUIRE(section == "STOCH" || section == "INDEP" || section == "ENDATA", SQLP_E_INVALID,
                    smps_where(path, l) + ": unsupported sto section " + section);
            keys.assign(l.tok.begin() + 1, l.tok.end());
            if (section == "STOCH" && !keys.empty()) p.sto_name = keys[0];
            continue;
        }
        if (section != "INDEP") continue;
        REQUIRE(keys.size() == 1, SQLP_E_INVALID,
                smps_where(path, l) + ": INDEP needs exactly one keyword (DISCRETE | NORMAL | UNIFORM; REPLACE mode only)");
        REQUIRE(l.tok.size() >= 4, SQLP_E_INVALID, smps_where(path, l) + ": INDEP line needs 'column row value value'");
        int kind = keys[0] == "DISCRETE" ? 0 : keys[0] == "NORMAL" ? 1 : keys[0] == "UNIFORM" ? 2 : -1;
        REQUIRE(kind >= 0, SQLP_E_INVALID, smps_where(path, l) + ": unsupported INDEP keyword " + keys[0]);
        auto key = std::make_pair(l.tok[0], l.tok[1]);
        double a = smps_number(l.tok[2], path, l), b = smps_number(l.tok[3], path, l);
        auto it = at.find(key);
        if (it == at.end()) {
            it = at.emplace(key, p.elems.size()).first;
            SmpsElement e;
            e.col = l.tok[0];
            e.row = l.tok[1];
            e.kind = kind;
            p.elems.push_back(e);
        }
        SmpsElement &e = p.elems[it->second];
        if (kind == 0) {
            if (e.kind != 0) { e.kind = 0; e.val.clear(); e.prob.clear(); }
            e.val.push_back(a);
            e.prob.push_back(b);
        } else {                               // a later line replaces the element (indep[pos] = ...)
            e.kind = kind;
            e.a = a;
            e.b = b;
            e.val.clear();
            e.prob.clear();
        }
    }
    p.has_sto = true;
}

// smps_prob.jl:14-102 (stage 2 of a two-period problem) followed by subprob.jl:15-69.
inline void smps_stage2(sqlp_smps &p)
{
    REQUIRE(p.periods.size() == 2, SQLP_E_UNSUPPORTED, "two-stage problems only (the tim file must list two periods)");
    auto c = p.col_of.find(p.periods[1][1]);
    auto r = p.row_of.find(p.periods[1][2]);
    REQUIRE(c != p.col_of.end(), SQLP_E_INVALID, "tim: unknown column " + p.periods[1][1]);
    REQUIRE(r != p.row_of.end(), SQLP_E_INVALID, "tim: unknown row " + p.periods[1][2]);
    p.c2 = c->second;
    p.r2 = r->second;
    REQUIRE(p.r2 >= 1, SQLP_E_INVALID, "tim: the second period starts at the objective row");
    p.n1 = p.c2;
    p.n2 = (int64_t)p.cols.size() - p.c2;
    p.m2 = (int64_t)p.rows.size() - p.r2;
    p.cost.assign(p.n2, 0.0);
    p.x_cost.assign(p.n1, 0.0);
    p.T_colptr.assign(p.n1 + 1, 0);
    p.W_colptr.assign(p.n2 + 1, 0);
    for (int64_t j = 0; j < (int64_t)p.cols.size(); ++j) {
        for (auto &kv : p.col_entries[j]) {             // rows ascending
            int64_t i = kv.first;
            double v = kv.second;
            if (i == 0) {
                (j < p.c2 ? p.x_cost[j] : p.cost[j - p.c2]) = v;
            } else if (i >= p.r2 && v != 0.0) {
                if (j < p.c2) { p.T_rowval.push_back(i - p.r2); p.T_nzval.push_back(v); }
                else { p.W_rowval.push_back(i - p.r2); p.W_nzval.push_back(v); }
            }
        }
        if (j < p.c2) p.T_colptr[j + 1] = (int64_t)p.T_rowval.size();
        else p.W_colptr[j - p.c2 + 1] = (int64_t)p.W_rowval.size();
    }
    for (int64_t i = 0; i < p.m2; ++i)
        if (p.rhs[p.r2 + i] != 0.0) { p.r_idx.push_back(i); p.r_val.push_back(p.rhs[p.r2 + i]); }
    for (const SmpsElement &e : p.elems) {
        auto rr = p.row_of.find(e.row);
        REQUIRE(rr != p.row_of.end(), SQLP_E_INVALID, "sto: unknown row " + e.row);
        int64_t row = rr->second - p.r2, col = -1;
        if (!(e.col == "RHS" || e.col == "rhs")) {       // subprob.jl:113
            auto cc = p.col_of.find(e.col);
            REQUIRE(cc != p.col_of.end(), SQLP_E_INVALID, "sto: unknown column " + e.col);
            col = cc->second;
        }
        REQUIRE(row >= 0 && col < p.n1, SQLP_E_INVALID,
                "sto: random element (" + e.col + ", " + e.row + ") is not in stage 2 / Tbar");
        p.pos_row.push_back((int32_t)row);
        p.pos_col.push_back((int32_t)col);
        if (e.kind == 0) p.max_outcomes = std::max<int64_t>(p.max_outcomes, (int64_t)e.val.size());
    }
}

}  // namespace
