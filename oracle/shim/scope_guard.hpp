/* TEST INFRASTRUCTURE — oracle build shim, never part of the product library.
 *
 * Stand-in for the `scope_guard` dependency the reference pulls through meson
 * (used as `MAKE_SCOPE_EXIT(name) { ... };` in groth16.cpp and multiexp.cpp).
 */
#ifndef KZP_ORACLE_SCOPE_GUARD_SHIM_H
#define KZP_ORACLE_SCOPE_GUARD_SHIM_H

#include <utility>

namespace kzp_shim
{
template <typename F>
struct ScopeExit
{
    F fn;
    explicit ScopeExit(F&& f)
        : fn(std::move(f))
    {
    }
    ScopeExit(ScopeExit&& o)
        : fn(std::move(o.fn))
    {
    }
    ScopeExit(ScopeExit const&)            = delete;
    ScopeExit& operator=(ScopeExit const&) = delete;
    ~ScopeExit() { fn(); }
};
struct ScopeExitMaker
{
    template <typename F>
    ScopeExit<F> operator+(F&& f) const
    {
        return ScopeExit<F>(std::move(f));
    }
};
} // namespace kzp_shim

#define MAKE_SCOPE_EXIT(name) auto name = ::kzp_shim::ScopeExitMaker{} + [&]()

#endif
