// Minimal stand-in for <Rcpp.h>: just the types and members rshim/bamsignals_shim.cpp AND the reference's own
// src/bamsignals.cpp use (the latter is compiled UNCHANGED against this header into oracle/_ref/, see oracle/Makefile),
// with the signatures of the real ones (Rcpp 1.0.x: Vector<INTSXP>, Matrix<INTSXP>, Vector<STRSXP>, Vector<VECSXP>,
// RObject, String, as<>, stop).
// TEST INFRASTRUCTURE: R and Rcpp cannot be installed offline, so tests/test_rshim_compiles.py compiles the shim
// against this header to check its C++ and, above all, every call into include/bamsignals_cuda.h (argument count,
// order and types).  The mock is functional enough to RUN the shim on plain C++ objects (see rshim_mock_driver.cpp).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

namespace Rcpp {

struct exception : std::runtime_error { using std::runtime_error::runtime_error; };
[[noreturn]] inline void stop(const std::string& msg) { throw exception(msg); }

class RObject;
class IntegerVector;
class IntegerMatrix;
class CharacterVector;
class List;

// one heap node per R object; attributes and S4 slots live in the same map, like ATTRIB()
struct Node {
    std::vector<int> ints;
    std::vector<std::string> strs;
    std::vector<std::shared_ptr<Node>> items;
    std::map<std::string, std::shared_ptr<Node>> attrs;
    std::vector<std::string> klass;
    int nrow = 0;
};
typedef std::shared_ptr<Node> SEXP;

class AttrProxy {
public:
    AttrProxy(SEXP owner, std::string name) : owner_(std::move(owner)), name_(std::move(name)) {}
    SEXP get() const { auto it = owner_->attrs.find(name_); if (it == owner_->attrs.end()) stop("no attribute/slot " + name_); return it->second; }
    template <class T> AttrProxy& operator=(const T& v) { owner_->attrs[name_] = v.sexp(); return *this; }
private:
    SEXP owner_;
    std::string name_;
};

class RObject {
public:
    RObject() : p_(std::make_shared<Node>()) {}
    RObject(SEXP p) : p_(std::move(p)) {}
    RObject(const AttrProxy& a) : p_(a.get()) {}
    bool inherits(const char* k) const { for (auto& c : p_->klass) if (c == k) return true; return false; }
    AttrProxy slot(const std::string& n) const { return AttrProxy(p_, n); }
    AttrProxy attr(const std::string& n) const { return AttrProxy(p_, n); }
    SEXP sexp() const { return p_; }
protected:
    SEXP p_;
};

class IntegerVector : public RObject {
public:
    typedef int* iterator;
    IntegerVector() {}
    // (two ints of slack capacity: the reference's Coverager::pileup increments array[0] of a zero-width region's EMPTY
    // vector, src/bamsignals.cpp:420-424 - in R that write lands behind the object's header; here it lands in the slack)
    explicit IntegerVector(R_xlen_t n) { p_->ints.reserve(size_t(n) + 2); p_->ints.assign(size_t(n), 0); }
    IntegerVector(SEXP p) : RObject(std::move(p)) {}
    R_xlen_t size() const { return R_xlen_t(p_->ints.size()); }
    R_xlen_t length() const { return size(); }
    int& operator[](R_xlen_t i) { return p_->ints[size_t(i)]; }
    const int& operator[](R_xlen_t i) const { return p_->ints[size_t(i)]; }
    iterator begin() { return p_->ints.data(); }
};

class IntegerMatrix : public RObject {
public:
    typedef int* iterator;
    IntegerMatrix(int nrow, R_xlen_t ncol) { p_->ints.reserve(size_t(nrow) * size_t(ncol) + 2); p_->ints.assign(size_t(nrow) * size_t(ncol), 0); p_->nrow = nrow; }
    iterator begin() { return p_->ints.data(); }
};

class StringProxy {
public:
    explicit StringProxy(std::string* s) : s_(s) {}
    operator std::string() const { return *s_; }
    const std::string& get() const { return *s_; }
private:
    std::string* s_;
};

// Rcpp::String: what indexing a CharacterVector yields once it leaves the vector (src/bamsignals.cpp:86-88,116-125)
class String {
public:
    String() {}
    String(const StringProxy& p) : s_(p.get()) {}
    String(std::string s) : s_(std::move(s)) {}
    operator std::string() const { return s_; }
    const char* get_cstring() const { return s_.c_str(); }
private:
    std::string s_;
};

class CharacterVector : public RObject {
public:
    CharacterVector() {}
    CharacterVector(SEXP p) : RObject(std::move(p)) {}
    R_xlen_t size() const { return R_xlen_t(p_->strs.size()); }
    R_xlen_t length() const { return size(); }
    StringProxy operator[](R_xlen_t i) const { return StringProxy(&p_->strs[size_t(i)]); }
    static CharacterVector create(const char* a, const char* b) { CharacterVector v; v.p_->strs = {a, b}; return v; }
};

class ItemProxy {
public:
    explicit ItemProxy(SEXP* slot) : slot_(slot) {}
    template <class T> ItemProxy& operator=(const T& v) { *slot_ = v.sexp(); return *this; }
    SEXP get() const { return *slot_; }
private:
    SEXP* slot_;
};

class List : public RObject {
public:
    explicit List(R_xlen_t n) { p_->items.assign(size_t(n), SEXP()); }
    List(SEXP p) : RObject(std::move(p)) {}
    R_xlen_t size() const { return R_xlen_t(p_->items.size()); }
    R_xlen_t length() const { return size(); }
    ItemProxy operator[](R_xlen_t i) { return ItemProxy(&p_->items[size_t(i)]); }
};

template <class T> struct AsImpl;
template <> struct AsImpl<RObject> { static RObject get(SEXP p) { return RObject(p); } };
template <> struct AsImpl<IntegerVector> { static IntegerVector get(SEXP p) { return IntegerVector(p); } };
template <> struct AsImpl<CharacterVector> { static CharacterVector get(SEXP p) { return CharacterVector(p); } };
template <class T> T as(const AttrProxy& a) { return AsImpl<T>::get(a.get()); }
template <class T> T as(const RObject& o) { return AsImpl<T>::get(o.sexp()); }
template <class T> T as(const StringProxy& s);
template <> inline std::string as<std::string>(const StringProxy& s) { return s.get(); }

}  // namespace Rcpp
