#include "kernel_registry.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
std::vector<b2_kernel_info*>& table() {
    static std::vector<b2_kernel_info*> t;
    return t;
}
// B200FFT_VARIANTS="kind:prec:n=variant,..." (tuning / experiments); unset -> variant 0 everywhere
int variant_override(int kind, int prec, int n) {
    const char* e = getenv("B200FFT_VARIANTS");
    if (!e) return 0;
    for (const char* p = e; *p;) {
        int k, pr, nn, v, used = 0;
        if (sscanf(p, "%d:%d:%d=%d%n", &k, &pr, &nn, &v, &used) == 4) {
            if (k == kind && pr == prec && nn == n) return v;
            p += used;
        } else {
            break;
        }
        if (*p == ',') ++p;
    }
    return 0;
}
}  // namespace

extern "C" void b2_register_kernel(const b2_kernel_info* k) {
    b2_kernel_info* m = const_cast<b2_kernel_info*>(k);
    int v = 0;
    for (const b2_kernel_info* o : table())
        if (o->kind == k->kind && o->prec == k->prec && o->n == k->n && o->inv == k->inv && o->ops == k->ops) ++v;
    m->variant = v;
    table().push_back(m);
}
extern "C" const b2_kernel_info* b2_find_kernel_variant(int kind, int prec, int n, int inv, int ops, int variant) {
    for (const b2_kernel_info* k : table())
        if (k->kind == kind && k->prec == prec && k->n == n && k->inv == inv && k->ops == ops && k->variant == variant)
            return k;
    return nullptr;
}
extern "C" const b2_kernel_info* b2_find_kernel(int kind, int prec, int n, int inv, int ops) {
    const int v = variant_override(kind, prec, n);
    const b2_kernel_info* k = v ? b2_find_kernel_variant(kind, prec, n, inv, ops, v) : nullptr;
    return k ? k : b2_find_kernel_variant(kind, prec, n, inv, ops, 0);
}
extern "C" int b2_kernel_count(void) { return (int)table().size(); }
extern "C" const b2_kernel_info* b2_kernel_at(int i) { return table()[i]; }
