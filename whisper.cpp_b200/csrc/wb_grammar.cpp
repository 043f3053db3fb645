// wb_grammar.cpp -- see wb_grammar.h.  Rule elements: a rule is a list of alternates separated by ALT and closed by END; an
// alternate is a sequence of RULE_REF and character classes; a class starts with CHAR (or CHAR_NOT for a negated class) and
// continues with CHAR_RNG_UPPER (upper bound of a range) / CHAR_ALT (another member).
#include "wb_grammar.h"
#include "wb_model.h"

namespace wb {
namespace {

using Elem = Grammar::Elem;
using Stack = Grammar::Stack;

inline bool ends_alternate(const Elem * e) { return e->type == WHISPER_GRETYPE_END || e->type == WHISPER_GRETYPE_ALT; }

// a token's text as code points (0-terminated) plus the UTF-8 sequence it leaves open; `remain < 0` flags an invalid sequence
struct Decoded { std::vector<uint32_t> cps; uint32_t value = 0; int remain = 0; };

Decoded decode_utf8(const char * s, uint32_t value, int remain) {            // whisper.cpp:5517-5571
    static const int seq_len[16] = { 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 2, 2, 3, 4 };
    Decoded d;
    const bool resuming = remain > 0;
    for (; *s && remain > 0; ++s, --remain) {                                 // finish the sequence a previous token left open
        const uint8_t b = (uint8_t) *s;
        if ((b >> 6) != 2) { d.cps.push_back(0); d.value = 0; d.remain = -1; return d; }
        value = (value << 6) + (b & 0x3F);
    }
    if (resuming && remain == 0) d.cps.push_back(value);
    while (*s) {
        const uint8_t first = (uint8_t) *s;
        remain = seq_len[first >> 4] - 1;
        if (remain < 0) { d.cps.assign(1, 0u); d.value = 0; d.remain = remain; return d; }   // stray continuation byte
        value = first & ((1u << (7 - remain)) - 1);
        ++s;
        for (; *s && remain > 0; ++s, --remain) value = (value << 6) + ((uint8_t) *s & 0x3F);
        if (remain == 0) d.cps.push_back(value);
    }
    d.cps.push_back(0);
    d.value = value; d.remain = remain;
    return d;
}

// does `chr` belong to the class starting at `pos`?  `after` receives the element following the class.
bool class_matches(const Elem * pos, uint32_t chr, const Elem ** after) {    // whisper.cpp:5584-5607
    const bool positive = pos->type == WHISPER_GRETYPE_CHAR;
    bool found = false;
    do {
        if (pos[1].type == WHISPER_GRETYPE_CHAR_RNG_UPPER) { found = found || (pos->value <= chr && chr <= pos[1].value); pos += 2; }
        else                                               { found = found || pos->value == chr; pos += 1; }
    } while (pos->type == WHISPER_GRETYPE_CHAR_ALT);
    if (after) *after = pos;
    return found == positive;
}

// could some completion of an open UTF-8 sequence belong to the class at `pos`?  (whisper.cpp:5611-5655)
bool class_matches_partial(const Elem * pos, uint32_t value, int remain) {
    const bool positive = pos->type == WHISPER_GRETYPE_CHAR;
    if (remain < 0 || (remain == 1 && value < 2)) return false;               // invalid, or an overlong 2-byte form
    uint32_t low = value << (remain * 6);
    const uint32_t high = low | ((1u << (remain * 6)) - 1);
    if (low == 0) { if (remain == 2) low = 1u << 11; else if (remain == 3) low = 1u << 16; }
    do {
        if (pos[1].type == WHISPER_GRETYPE_CHAR_RNG_UPPER) { if (pos->value <= high && low <= pos[1].value) return positive; pos += 2; }
        else                                               { if (low <= pos->value && pos->value <= high) return positive; pos += 1; }
    } while (pos->type == WHISPER_GRETYPE_CHAR_ALT);
    return !positive;
}

// expand rule references on top of `stack` until every resulting stack is empty or has a character class on top (5660-5711)
void expand(const std::vector<std::vector<Elem>> & rules, const Stack & stack, std::vector<Stack> & out) {
    if (stack.empty()) { out.emplace_back(); return; }
    const Elem * top = stack.back();
    if (top->type == WHISPER_GRETYPE_CHAR || top->type == WHISPER_GRETYPE_CHAR_NOT) { out.push_back(stack); return; }
    if (top->type != WHISPER_GRETYPE_RULE_REF) return;                         // malformed grammar: a stack never rests on END/ALT/RNG/CHAR_ALT
    const Elem * alt = rules[(size_t) top->value].data();
    for (;;) {                                                                 // one new stack per alternate of the referenced rule
        Stack next(stack.begin(), stack.end() - 1);
        if (!ends_alternate(top + 1)) next.push_back(top + 1);                 // what follows the reference
        if (!ends_alternate(alt)) next.push_back(alt);                         // the alternate itself (may be empty)
        expand(rules, next, out);
        while (!ends_alternate(alt)) ++alt;
        if (alt->type != WHISPER_GRETYPE_ALT) break;
        ++alt;
    }
}

struct Cand { whisper_token id; const uint32_t * cp; uint32_t value; int remain; };

std::vector<Cand> rejected_by_all(const std::vector<std::vector<Elem>> & rules, const std::vector<Stack> & stacks, const std::vector<Cand> & cands);

// candidates that a single parse cannot consume completely (5750-5799)
std::vector<Cand> rejected_by(const std::vector<std::vector<Elem>> & rules, const Stack & stack, const std::vector<Cand> & cands) {
    std::vector<Cand> rejects;
    if (stack.empty()) {                                                       // the parse is complete: only the empty remainder fits
        for (const Cand & c : cands) if (*c.cp != 0 || c.remain != 0) rejects.push_back(c);
        return rejects;
    }
    const Elem * top = stack.back();
    std::vector<Cand> rest;
    for (const Cand & c : cands) {
        if (*c.cp == 0) { if (c.remain != 0 && !class_matches_partial(top, c.value, c.remain)) rejects.push_back(c); }
        else if (class_matches(top, *c.cp, nullptr)) rest.push_back({ c.id, c.cp + 1, c.value, c.remain });
        else rejects.push_back(c);
    }
    const Elem * after = nullptr;
    class_matches(top, 0, &after);
    Stack below(stack.begin(), stack.end() - 1);
    if (!ends_alternate(after)) below.push_back(after);
    std::vector<Stack> next;
    expand(rules, below, next);
    for (const Cand & c : rejected_by_all(rules, next, rest)) rejects.push_back({ c.id, c.cp - 1, c.value, c.remain });
    return rejects;
}

// candidates that NO live parse can consume (5801-5815)
std::vector<Cand> rejected_by_all(const std::vector<std::vector<Elem>> & rules, const std::vector<Stack> & stacks, const std::vector<Cand> & cands) {
    if (cands.empty() || stacks.empty()) return {};
    std::vector<Cand> rejects = rejected_by(rules, stacks.front(), cands);
    for (size_t i = 1; i < stacks.size(); ++i) rejects = rejected_by(rules, stacks[i], rejects);
    return rejects;
}

} // namespace

Grammar grammar_init(const whisper_grammar_element * const * rules, size_t n_rules, size_t i_start_rule) {
    Grammar g;
    if (!rules || n_rules == 0 || i_start_rule >= n_rules) return g;
    auto table = std::make_shared<std::vector<std::vector<Elem>>>(n_rules);
    for (size_t i = 0; i < n_rules; ++i) {
        for (const Elem * e = rules[i]; e->type != WHISPER_GRETYPE_END; ++e) (*table)[i].push_back(*e);
        (*table)[i].push_back({ WHISPER_GRETYPE_END, 0 });
    }
    g.rules = table;
    const Elem * alt = (*table)[i_start_rule].data();              // all stack entries point into the shared copy
    for (;;) {
        Grammar::Stack st;
        if (!ends_alternate(alt)) st.push_back(alt);
        expand(*g.rules, st, g.stacks);
        while (!ends_alternate(alt)) ++alt;
        if (alt->type != WHISPER_GRETYPE_ALT) break;
        ++alt;
    }
    return g;
}

void grammar_penalize(const Vocab & vocab, const Grammar & g, float penalty, std::vector<float> & logits) {
    if (!g.active()) return;
    std::vector<Decoded> decoded;
    std::vector<Cand> cands;
    decoded.reserve((size_t) vocab.token_eot);
    for (whisper_token id = 0; id < vocab.token_eot; ++id) {
        auto it = vocab.id_to_token.find(id);
        if (it == vocab.id_to_token.end() || it->second.empty()) continue;
        decoded.push_back(decode_utf8(it->second.c_str(), g.partial_value, g.partial_remain));
    }
    size_t k = 0;
    for (whisper_token id = 0; id < vocab.token_eot; ++id) {                  // second pass: `decoded` no longer reallocates
        auto it = vocab.id_to_token.find(id);
        if (it == vocab.id_to_token.end() || it->second.empty()) continue;
        const Decoded & d = decoded[k++];
        cands.push_back({ id, d.cps.data(), d.value, d.remain });
    }
    for (const Cand & c : rejected_by_all(*g.rules, g.stacks, cands)) logits[c.id] -= penalty;
}

void grammar_accept_token(const Vocab & vocab, Grammar & g, whisper_token token) {
    if (!g.active()) return;
    auto it = vocab.id_to_token.find(token);
    if (it == vocab.id_to_token.end()) return;
    const std::string & text = it->second;
    if (text.rfind("[_", 0) == 0) return;                                      // special tokens are invisible to the grammar
    const Decoded d = decode_utf8(text.c_str(), g.partial_value, g.partial_remain);
    for (size_t i = 0; i + 1 < d.cps.size(); ++i) {                            // every code point but the terminating 0
        std::vector<Grammar::Stack> next;
        for (const Grammar::Stack & st : g.stacks) {
            if (st.empty()) continue;
            const Elem * after = nullptr;
            if (!class_matches(st.back(), d.cps[i], &after)) continue;
            Grammar::Stack ns(st.begin(), st.end() - 1);
            if (!ends_alternate(after)) ns.push_back(after);
            expand(*g.rules, ns, next);
        }
        g.stacks = std::move(next);
    }
    g.partial_value = d.value; g.partial_remain = d.remain;
}

} // namespace wb
