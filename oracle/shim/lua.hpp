// lua.hpp -- stand-in for the Lua C API (test infrastructure, part of oracle/; see shim/opencv2/opencv.hpp).
//
// The reference reaches its two CNNs through an in-process Lua VM (core/lua_calls.h): it pushes numbers into tables,
// calls the global functions "forward" / "backward" / "loadModel" / ... with lua_pcall and reads numbers back.  Neither
// Lua nor Torch7 exists in this image, and the CNNs are out of scope anyway: north_star replaces the score CNN by the
// closed-form soft-inlier score, and the coordinate CNN's OUTPUT is the engine's input.  This header implements just
// enough of the C API (a value stack with numbers, strings, flat number tables and named functions) for
// core/lua_calls.h to compile and run UNMODIFIED; a call of a global function is dispatched to a C++ handler the
// harness registers per lua_State role (oracle/ref_harness):
//     score state :  forward(count, maps)            -> count scores          (soft-inlier closed form)
//                    backward(count, maps, outGrads) -> table of count*1600 gradients, laid out [c][row][col] like the
//                                                       maps were pushed -- lua_calls.h:326-338 reads it back transposed
//     coord state :  forward(count, patches)         -> table of count*3 scene coordinates in metres
//                    backward(count, loss, patches, dLoss) -> captured by the harness (the final gradient)
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#define LUA_MULTRET (-1)

struct shim_lua_value {
    enum Kind { NIL, NUMBER, STRING, TABLE, FUNCTION } kind = NIL;
    double num = 0;
    std::string str;
    std::shared_ptr<std::vector<double>> tab;   // 1-based flat number table: tab[i-1]
};

struct lua_State {
    std::vector<shim_lua_value> stack;
    std::string script;   // file passed to luaL_loadfile: identifies the role of the state
    void* user = nullptr;
};

// implemented by the harness: called for lua_pcall of a named global; arguments are the top nargs stack values
// (already popped into `args`), results are appended to `results`
extern "C++" void shim_lua_dispatch(lua_State* L, const std::string& fn, std::vector<shim_lua_value>& args, int nresults,
                                    std::vector<shim_lua_value>& results);

inline lua_State* luaL_newstate() { return new lua_State(); }
inline void luaL_openlibs(lua_State*) {}
inline void lua_close(lua_State* L) { delete L; }
inline int lua_gettop(lua_State* L) { return (int)L->stack.size(); }
inline shim_lua_value& shim_lua_at(lua_State* L, int idx) { return idx > 0 ? L->stack[idx - 1] : L->stack[L->stack.size() + idx]; }
inline void lua_pop(lua_State* L, int n) { L->stack.resize(L->stack.size() - n); }
inline void lua_pushnumber(lua_State* L, double v) { shim_lua_value x; x.kind = shim_lua_value::NUMBER; x.num = v; L->stack.push_back(std::move(x)); }
inline void lua_pushinteger(lua_State* L, long long v) { lua_pushnumber(L, (double)v); }
inline void lua_pushstring(lua_State* L, const char* s) { shim_lua_value x; x.kind = shim_lua_value::STRING; x.str = s; L->stack.push_back(std::move(x)); }
inline void lua_createtable(lua_State* L, int narr, int) {
    shim_lua_value x; x.kind = shim_lua_value::TABLE; x.tab = std::make_shared<std::vector<double>>();
    x.tab->reserve(narr > 0 ? narr : 0);
    L->stack.push_back(std::move(x));
}
inline void lua_rawseti(lua_State* L, int idx, int n) {   // t[n] = top; pops the value (idx is resolved with the value still on the stack)
    std::shared_ptr<std::vector<double>> t = shim_lua_at(L, idx).tab;
    const double v = L->stack.back().num;
    L->stack.pop_back();
    if ((int)t->size() < n) t->resize(n, 0.0);
    (*t)[n - 1] = v;
}
inline void lua_gettable(lua_State* L, int idx) {   // key = top (popped); pushes t[key] (idx is resolved with the key still on the stack)
    const shim_lua_value t = shim_lua_at(L, idx);
    const int key = (int)L->stack.back().num;
    L->stack.pop_back();
    double v = 0;
    if (t.kind == shim_lua_value::TABLE && key >= 1 && key <= (int)t.tab->size()) v = (*t.tab)[key - 1];
    lua_pushnumber(L, v);
}
inline double lua_tonumber(lua_State* L, int idx) { return shim_lua_at(L, idx).num; }
inline const char* lua_tostring(lua_State* L, int idx) { return shim_lua_at(L, idx).str.c_str(); }
inline void lua_getglobal(lua_State* L, const char* name) { shim_lua_value x; x.kind = shim_lua_value::FUNCTION; x.str = name; L->stack.push_back(std::move(x)); }
inline int luaL_loadfile(lua_State* L, const char* filename) {
    L->script = filename;
    lua_getglobal(L, "__chunk__");
    return 0;
}
inline int lua_pcall(lua_State* L, int nargs, int nresults, int) {
    std::vector<shim_lua_value> args(L->stack.end() - nargs, L->stack.end());
    L->stack.resize(L->stack.size() - nargs);
    const std::string fn = L->stack.back().str;
    L->stack.pop_back();
    std::vector<shim_lua_value> results;
    if (fn != "__chunk__") shim_lua_dispatch(L, fn, args, nresults, results);
    if (nresults != LUA_MULTRET) results.resize(nresults);
    for (auto& r : results) L->stack.push_back(r);
    return 0;
}
