#include "kernel_registry.h"
#include <vector>

namespace {
std::vector<const b2_kernel_info*>& table() {
    static std::vector<const b2_kernel_info*> t;
    return t;
}
}  // namespace

extern "C" void b2_register_kernel(const b2_kernel_info* k) { table().push_back(k); }
extern "C" const b2_kernel_info* b2_find_kernel(int kind, int prec, int n, int inv, int ops) {
    for (const b2_kernel_info* k : table())
        if (k->kind == kind && k->prec == prec && k->n == n && k->inv == inv && k->ops == ops) return k;
    return nullptr;
}
extern "C" int b2_kernel_count(void) { return (int)table().size(); }
extern "C" const b2_kernel_info* b2_kernel_at(int i) { return table()[i]; }
