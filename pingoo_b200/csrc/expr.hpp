// Parser for the rule expression language: the CEL subset Pingoo documents in
// docs/rules.md:35-76 and compiles with `bel::Program::compile`
// (reference rules/rules.rs:45-53).  Produces a plain AST; typing and lowering
// to device atoms happen in lower.cpp.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pgw {

struct Expr;
using ExprP = std::unique_ptr<Expr>;

struct Expr {
    enum Kind : uint8_t {
        LIT_NULL, LIT_BOOL, LIT_INT, LIT_UINT, LIT_FLOAT, LIT_STR, LIT_BYTES,
        IDENT,     // name
        MEMBER,    // kids[0].name
        INDEX,     // kids[0][kids[1]]
        CALL,      // name(kids...)
        METHOD,    // kids[0].name(kids[1..])
        UNARY,     // op kids[0]       op in { '!', '-' }
        BINARY,    // kids[0] op kids[1]
        TERNARY,   // kids[0] ? kids[1] : kids[2]
        LIST,      // [kids...]
        MAP        // {k0: v0, k1: v1, ...} as kids pairs
    } kind;
    std::string name;  // IDENT/MEMBER/CALL/METHOD name, LIT_STR/LIT_BYTES payload
    int64_t ival = 0;
    double fval = 0;
    bool bval = false;
    enum Op : uint8_t { OP_NONE, OP_NOT, OP_NEG, OP_OR, OP_AND, OP_EQ, OP_NE, OP_LT, OP_LE, OP_GT, OP_GE, OP_IN,
                        OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD } op = OP_NONE;
    std::vector<ExprP> kids;
    size_t pos = 0;  // byte offset in the source, for messages
};

// Returns nullptr and sets `err` on a syntax error.
ExprP parse_expression(const std::string& src, std::string& err);

// Names of every function/operator the expression references, in the style of
// bel::Program::references().functions() (reference rules/rules.rs:65-69):
// method and call names, plus "@in" for the `in` operator.
void collect_functions(const Expr& e, std::vector<std::string>& out);

}  // namespace pgw
