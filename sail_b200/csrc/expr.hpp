// expr.hpp -- physical expressions: parsing from the JSON spec and DataFusion/arrow-rs type rules.
//
// Mirrors what Sail hands DataFusion after planning: BinaryExpr / Literal / Column / CastExpr /
// CaseExpr / InListExpr / LikeExpr / NotExpr / IsNull / ScalarFunction(date_part, substr)
// (crates/sail-plan/src/function/scalar/math.rs:48-181,580-583; predicate.rs:103-125).
// Result types follow arrow-arith 58.
#pragma once
#include <algorithm>

#include "common.hpp"

namespace sg {

struct Expr;
using ExprPtr = std::shared_ptr<Expr>;

struct Expr {
  enum Kind { Col, Lit, Bin, Not, Neg, IsNull, IsNotNull, Cast, Case, Like, DatePart, Substr } kind = Col;
  DataType type;
  bool nullable = false;
  // Col
  int col = -1;
  // Lit
  bool lit_null = false;
  i128 lit_i = 0;        // ints, decimals (unscaled), dates, bools
  double lit_f = 0.0;
  std::string lit_s;
  // Bin / Like / DatePart
  std::string op;        // "+", "=", "and", ... ; LIKE pattern ; date part
  bool negated = false;
  long long sub_start = 1, sub_len = -1;     // Substr: 1-based first character, character count (-1: to the end)
  std::vector<ExprPtr> args;   // Bin: l, r ; Case: w0,t0,w1,t1,...,[else]
  bool has_else = false;

  std::string key() const;     // structural key for common-subexpression elimination
};

inline std::string i128_str(i128 v) {
  if (v == 0) return "0";
  bool neg = v < 0; u128 u = neg ? (u128)0 - (u128)v : (u128)v; std::string s;
  while (u) { s += (char)('0' + (int)(u % 10)); u /= 10; }
  if (neg) s += '-';
  std::reverse(s.begin(), s.end());
  return s;
}

inline std::string Expr::key() const {
  switch (kind) {
    case Col: return "c" + std::to_string(col);
    case Lit: return "l[" + type.str() + ":" + (lit_null ? "null" : type.is_string() ? lit_s : type.is_float() ? std::to_string(lit_f) : i128_str(lit_i)) + "]";
    default: {
      std::string s = std::to_string((int)kind) + "(" + op + (negated ? "!" : "") + ":" + type.str();
      for (auto& a : args) s += "," + a->key();
      return s + ")";
    }
  }
}

inline DataType int_as_decimal(const DataType& t) {
  switch (t.id) {
    case TypeId::Int8: case TypeId::UInt8: return Dec(3, 0);
    case TypeId::Int16: case TypeId::UInt16: return Dec(5, 0);
    case TypeId::Int32: case TypeId::UInt32: return Dec(10, 0);
    default: return Dec(20, 0);
  }
}

inline DataType decimal_result(const std::string& op, const DataType& a, const DataType& b) {
  int p1 = a.precision, s1 = a.scale, p2 = b.precision, s2 = b.scale;
  if (op == "+" || op == "-") { int s = std::max(s1, s2); return Dec(std::min(38, std::max(p1 - s1, p2 - s2) + s + 1), s); }
  if (op == "*") return Dec(std::min(38, p1 + p2 + 1), s1 + s2);
  if (op == "/") { int s = std::min(38, s1 + 4); return Dec(std::min(38, p1 - s1 + s2 + s), s); }
  if (op == "%") { int s = std::max(s1, s2); return Dec(std::min(38, std::min(p1 - s1, p2 - s2) + s), s); }
  fail(SAILGPU_ERR_INVALID, "bad decimal op " + op);
}

inline ExprPtr make_cast(ExprPtr e, const DataType& to) {
  if (e->type == to) return e;
  auto c = std::make_shared<Expr>();
  c->kind = Expr::Cast; c->type = to; c->nullable = e->nullable; c->args = {e};
  return c;
}

inline bool is_arith(const std::string& op) { return op == "+" || op == "-" || op == "*" || op == "/" || op == "%"; }
inline bool is_cmp(const std::string& op) { return op == "=" || op == "!=" || op == "<" || op == "<=" || op == ">" || op == ">="; }

// numeric coercion for operands that arrive with different types (plans normally arrive coerced)
inline void coerce_numeric(ExprPtr& l, ExprPtr& r, bool for_compare) {
  const DataType a = l->type, b = r->type;
  if (a == b) return;
  if (a.is_decimal() && b.is_int()) { r = make_cast(r, int_as_decimal(b)); }
  else if (b.is_decimal() && a.is_int()) { l = make_cast(l, int_as_decimal(a)); }
  else if (a.is_float() || b.is_float()) { l = make_cast(l, T(TypeId::Float64)); r = make_cast(r, T(TypeId::Float64)); }
  else if (a.is_int() && b.is_int()) { l = make_cast(l, T(TypeId::Int64)); r = make_cast(r, T(TypeId::Int64)); }
  if (for_compare && l->type.is_decimal() && r->type.is_decimal() && l->type != r->type) {
    // compare at the wider scale / integer-digit count
    int s = std::max(l->type.scale, r->type.scale);
    int ip = std::max(l->type.precision - l->type.scale, r->type.precision - r->type.scale);
    DataType t = Dec(std::min(38, ip + s), s);
    l = make_cast(l, t); r = make_cast(r, t);
  }
}

ExprPtr parse_expr(const Json& j, const Schema& in);

inline ExprPtr parse_literal(const Json& j) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Lit;
  e->type = parse_type(j.at("type").as_str());
  const Json& v = j.at("lit");
  if (v.is_null()) { e->lit_null = true; e->nullable = true; return e; }
  if (e->type.is_string()) e->lit_s = v.as_str();
  else if (e->type.is_float()) e->lit_f = v.as_double();
  else if (e->type.id == TypeId::Bool) e->lit_i = v.kind == Json::Bool ? (v.b ? 1 : 0) : (v.as_int() != 0);
  else e->lit_i = parse_i128(v.s);
  return e;
}

inline ExprPtr make_bin(const std::string& op, ExprPtr l, ExprPtr r) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Bin; e->op = op;
  if (is_arith(op)) {
    coerce_numeric(l, r, false);
    if (l->type.is_decimal() && r->type.is_decimal()) e->type = decimal_result(op, l->type, r->type);
    else {
      SG_CHECK(l->type == r->type, SAILGPU_ERR_UNSUPPORTED, "arithmetic on " + l->type.str() + " and " + r->type.str());
      SG_CHECK(l->type.is_int() || l->type.is_float(), SAILGPU_ERR_UNSUPPORTED, "arithmetic on " + l->type.str());
      e->type = l->type;
    }
  } else if (is_cmp(op)) {
    if (l->type.is_string() && r->type.is_string()) {
      SG_CHECK(op == "=" || op == "!=", SAILGPU_ERR_UNSUPPORTED, "ordering comparison on strings is not supported on the GPU path yet");
    } else {
      coerce_numeric(l, r, true);
      SG_CHECK(l->type == r->type, SAILGPU_ERR_UNSUPPORTED, "comparison of " + l->type.str() + " and " + r->type.str());
    }
    e->type = T(TypeId::Bool);
  } else if (op == "and" || op == "or") {
    SG_CHECK(l->type.id == TypeId::Bool && r->type.id == TypeId::Bool, SAILGPU_ERR_INVALID, "AND/OR need boolean operands");
    e->type = T(TypeId::Bool);
  } else {
    fail(SAILGPU_ERR_UNSUPPORTED, "binary operator '" + op + "'");
  }
  e->nullable = l->nullable || r->nullable;
  e->args = {l, r};
  return e;
}

inline ExprPtr parse_expr(const Json& j, const Schema& in) {
  SG_CHECK(j.kind == Json::Obj, SAILGPU_ERR_INVALID, "spec: expression must be an object");
  if (j.has("col")) {
    auto e = std::make_shared<Expr>();
    int c = (int)j.at("col").as_int();
    SG_CHECK(c >= 0 && c < (int)in.size(), SAILGPU_ERR_INVALID, "spec: column index " + std::to_string(c) + " out of range");
    e->kind = Expr::Col; e->col = c; e->type = in[c].type; e->nullable = in[c].nullable;
    return e;
  }
  if (j.has("lit")) return parse_literal(j);
  if (j.has("op")) return make_bin(j.at("op").as_str(), parse_expr(j.at("l"), in), parse_expr(j.at("r"), in));
  auto e = std::make_shared<Expr>();
  if (j.has("not")) {
    e->kind = Expr::Not; e->args = {parse_expr(j.at("not"), in)}; e->type = T(TypeId::Bool); e->nullable = e->args[0]->nullable;
    SG_CHECK(e->args[0]->type.id == TypeId::Bool, SAILGPU_ERR_INVALID, "NOT needs a boolean operand");
    return e;
  }
  if (j.has("neg")) {
    e->kind = Expr::Neg; e->args = {parse_expr(j.at("neg"), in)}; e->type = e->args[0]->type; e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.has("is_null") || j.has("is_not_null")) {
    const bool isn = j.has("is_null");
    e->kind = isn ? Expr::IsNull : Expr::IsNotNull;
    e->args = {parse_expr(j.at(isn ? "is_null" : "is_not_null"), in)}; e->type = T(TypeId::Bool);
    return e;
  }
  if (j.has("cast")) return make_cast(parse_expr(j.at("cast"), in), parse_type(j.at("to").as_str()));
  if (j.has("case")) {
    e->kind = Expr::Case;
    std::vector<ExprPtr> thens;
    for (auto& br : j.at("case").a) {
      SG_CHECK(br.kind == Json::Arr && br.a.size() == 2, SAILGPU_ERR_INVALID, "spec: CASE branch must be [when, then]");
      e->args.push_back(parse_expr(br.a[0], in));
      e->args.push_back(parse_expr(br.a[1], in));
      thens.push_back(e->args.back());
    }
    const Json* el = j.find("else");
    ExprPtr els;
    if (el && !el->is_null()) { els = parse_expr(*el, in); thens.push_back(els); }
    DataType rt = thens[0]->type;
    bool alldec = true; for (auto& t : thens) alldec &= t->type.is_decimal();
    if (alldec) {
      int s = 0, ip = 0;
      for (auto& t : thens) { s = std::max(s, t->type.scale); ip = std::max(ip, t->type.precision - t->type.scale); }
      rt = Dec(std::min(38, ip + s), s);
    }
    for (size_t i = 1; i < e->args.size(); i += 2) e->args[i] = make_cast(e->args[i], rt);
    e->nullable = !els;
    for (size_t i = 1; i < e->args.size(); i += 2) e->nullable |= e->args[i]->nullable;
    if (els) { els = make_cast(els, rt); e->nullable |= els->nullable; e->args.push_back(els); e->has_else = true; }
    e->type = rt;
    return e;
  }
  if (j.has("in")) {
    // InListExpr over literals == OR of equalities (NULL semantics identical for non-null lists)
    ExprPtr x = parse_expr(j.at("in"), in), acc;
    for (auto& l : j.at("set").a) {
      ExprPtr eq = make_bin("=", x, parse_literal(l));
      acc = acc ? make_bin("or", acc, eq) : eq;
    }
    SG_CHECK((bool)acc, SAILGPU_ERR_INVALID, "spec: empty IN list");
    const Json* neg = j.find("negated");
    if (neg && neg->kind == Json::Bool && neg->b) {
      auto n = std::make_shared<Expr>();
      n->kind = Expr::Not; n->args = {acc}; n->type = T(TypeId::Bool); n->nullable = acc->nullable;
      return n;
    }
    return acc;
  }
  if (j.has("like")) {
    e->kind = Expr::Like; e->args = {parse_expr(j.at("like"), in)}; e->op = j.at("pattern").as_str();
    const Json* neg = j.find("negated"); e->negated = neg && neg->kind == Json::Bool && neg->b;
    SG_CHECK(e->args[0]->type.is_string(), SAILGPU_ERR_INVALID, "LIKE needs a string operand");
    e->type = T(TypeId::Bool); e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.has("fn")) {
    const std::string fn = j.at("fn").as_str();
    if (fn == "date_part") {
      e->kind = Expr::DatePart; e->op = j.at("part").as_str();
      std::transform(e->op.begin(), e->op.end(), e->op.begin(), ::tolower);
      SG_CHECK(e->op == "year" || e->op == "month" || e->op == "day", SAILGPU_ERR_UNSUPPORTED, "date_part('" + e->op + "')");
      e->args = {parse_expr(j.at("args").a.at(0), in)};
      SG_CHECK(e->args[0]->type.id == TypeId::Date32, SAILGPU_ERR_UNSUPPORTED, "date_part on " + e->args[0]->type.str());
      e->type = T(TypeId::Int32); e->nullable = e->args[0]->nullable;
      return e;
    }
    if (fn == "substr") {      // substr(str, start[, length]) with literal positions (Spark / DataFusion character semantics)
      e->kind = Expr::Substr;
      e->args = {parse_expr(j.at("args").a.at(0), in)};
      SG_CHECK(e->args[0]->type.is_string(), SAILGPU_ERR_INVALID, "substr needs a string operand");
      e->sub_start = j.at("start").as_int();
      const Json* ln = j.find("length");
      e->sub_len = ln && !ln->is_null() ? ln->as_int() : -1;
      SG_CHECK(e->sub_start >= 1, SAILGPU_ERR_UNSUPPORTED, "substr with a start position below 1 is not supported on the GPU path yet");
      SG_CHECK(!(ln && !ln->is_null()) || e->sub_len >= 0, SAILGPU_ERR_INVALID, "substr with a negative length");
      e->op = "substr:" + std::to_string(e->sub_start) + ":" + std::to_string(e->sub_len);
      e->type = e->args[0]->type; e->nullable = e->args[0]->nullable;
      return e;
    }
    fail(SAILGPU_ERR_UNSUPPORTED, "scalar function '" + fn + "' is not implemented on the GPU path");
  }
  fail(SAILGPU_ERR_INVALID, "spec: unrecognised expression object");
}

// ---- aggregate typing (DataFusion UDAFs) ------------------------------------
struct AggTypes { std::vector<DataType> state; DataType final_type; };
inline AggTypes agg_types(const std::string& fn, const DataType& in) {
  AggTypes t;
  if (fn == "count") { t.state = {T(TypeId::Int64)}; t.final_type = T(TypeId::Int64); return t; }
  if (fn == "min" || fn == "max") { t.state = {in}; t.final_type = in; return t; }
  if (fn == "sum") {
    DataType s = in.is_decimal() ? Dec(std::min(38, in.precision + 10), in.scale)
               : in.is_float() ? T(TypeId::Float64) : in.is_unsigned_int() ? T(TypeId::UInt64) : T(TypeId::Int64);
    t.state = {s}; t.final_type = s; return t;
  }
  if (fn == "avg") {
    if (in.is_decimal()) {
      t.state = {T(TypeId::UInt64), Dec(std::min(38, in.precision + 10), in.scale)};
      t.final_type = Dec(std::min(38, in.precision + 4), std::min(38, in.scale + 4));
    } else {
      t.state = {T(TypeId::UInt64), T(TypeId::Float64)};
      t.final_type = T(TypeId::Float64);
    }
    return t;
  }
  fail(SAILGPU_ERR_UNSUPPORTED, "aggregate function '" + fn + "'");
}

}  // namespace sg
