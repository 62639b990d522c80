// Host-only exercise of sym::Values / Rot3 / Pose3 of include/sym/sym.h (no GPU): prints named number lists that
// tests/test_sym_values_cpu.py recomputes with symforce_b200/geo.py.
#include <cstdio>
#include <string>
#include <vector>

#include <sym/sym.h>

static void Print(const char* name, const double* p, size_t n) {
  std::printf("%s", name);
  for (size_t i = 0; i < n; ++i) std::printf(" %.17g", p[i]);
  std::printf("\n");
}
static void Print(const char* name, const std::vector<double>& v) { Print(name, v.data(), v.size()); }

int main() {
  const double eps = sym::kDefaultEpsilond;
  const sym::Pose3d a(sym::Rot3d::FromTangent(sym::Vector3d(0.3, -0.2, 0.5), eps), sym::Vector3d(1.0, 2.0, 3.0));
  const sym::Pose3d b(sym::Rot3d::FromTangent(sym::Vector3d(-0.7, 0.1, 0.9), eps), sym::Vector3d(-0.5, 0.25, 4.0));
  Print("a", a.Data().data(), 7);
  Print("b", b.Data().data(), 7);
  Print("a_inv", a.Inverse().Data().data(), 7);
  Print("ab", a.Compose(b).Data().data(), 7);
  Print("a_local_b", a.LocalCoordinates(b, eps).data(), 6);
  Print("rot_tangent", b.Rotation().ToTangent(eps).data(), 3);

  sym::Valuesd v;
  v.Set('p', a);
  v.Set('r', b.Rotation());
  v.Set({'x', 1}, sym::Vector3d(1.0, -1.0, 0.5));
  v.Set('s', 2.5);
  sym::Valuesd w = v;
  w.Set('p', b);
  w.Set('r', a.Rotation());
  w.Set({'x', 1}, sym::Vector3d(0.0, 4.0, 0.25));
  w.Set('s', -1.0);

  // Keys / CreateIndex in storage order
  std::string order;
  for (const sym::Key& k : v.Keys()) order += k.str() + ",";
  std::printf("keys %s\n", order.c_str());
  const sym::index_t all = v.CreateIndex(true);
  std::printf("index %d %d %zu\n", all.storage_dim, all.tangent_dim, all.entries.size());

  // LocalCoordinates then Retract by it reproduces `w` (up to the sign of quaternions)
  const std::vector<sym::Key> keys = {'p', 'r', {'x', 1}, 's'};
  const sym::index_t idx = v.CreateIndex(keys);
  const std::vector<double> delta = v.LocalCoordinates(w, idx, eps);
  Print("delta", delta);
  sym::Valuesd u = v;
  u.Retract(idx, delta.data(), eps);
  Print("retracted", u.Data());
  Print("target", w.Data());

  // Update copies only the indexed entries
  sym::Valuesd c = v;
  c.Update(v.CreateIndex(std::vector<sym::Key>{'r', 's'}), w);
  Print("updated", c.Data());

  // Remove + Cleanup compacts; UpdateOrSet appends what is missing
  sym::Valuesd d = v;
  const bool removed = d.Remove('r'), again = d.Remove('r');
  const size_t freed = d.Cleanup();
  std::printf("remove %d %d %zu %zu %d\n", (int)removed, (int)again, freed, d.Data().size(), (int)d.Has('r'));
  Print("compacted", d.Data());
  d.UpdateOrSet(w.CreateIndex(std::vector<sym::Key>{'r', 's'}), w);
  Print("update_or_set", d.Data());
  std::printf("entry %d\n", (int)d.MaybeIndexEntryAt('r').has_value() + 2 * (int)d.MaybeIndexEntryAt('q').has_value());
  Print("at_entry", d.At<sym::Vector3d>(d.IndexEntryAt({'x', 1})).data(), 3);
  d.Set(d.IndexEntryAt('s'), 7.0);
  std::printf("set_entry %.17g\n", d.At<double>('s'));
  bool threw = false;
  try {
    d.SetNew('s', 1.0);
  } catch (const std::runtime_error&) {
    threw = true;
  }
  d.RemoveAll();
  std::printf("misc %d %d\n", (int)threw, (int)d.Empty());

  // Factor::Jacobian (factor.h:195-196) lowers a generated function onto the same device kind as Factor::Hessian;
  // keys_to_optimize defaults to the linearized arguments; host functors are rejected (no CPU fallback)
  const std::vector<sym::Key> bkeys = {{'P', 0}, {'P', 1}, {'T', 0}, {'I', 0}, 'e'};
  const sym::Factord fj = sym::Factord::Jacobian(sym::BetweenFactorPose3<double>, bkeys);
  const sym::Factord fh = sym::Factord::Hessian(sym::BetweenFactorPose3<double>, bkeys, {{'P', 0}, {'P', 1}});
  bool lambda_rejected = false;
  try {
    (void)sym::Factord::Jacobian([](double, double*, double*) {}, {'s'});
  } catch (const std::runtime_error&) {
    lambda_rejected = true;
  }
  bool wrong_keys_rejected = false;
  try {
    (void)sym::Factord::Jacobian(sym::BetweenFactorPose3<double>, bkeys, {{'P', 1}, {'P', 0}});
  } catch (const std::runtime_error&) {
    wrong_keys_rejected = true;
  }
  std::printf("factor_jacobian %d %d %zu %zu %d %d\n", (int)(fj.Kind() == fh.Kind()), (int)(fj.OptimizedKeys() == fh.OptimizedKeys()),
              fj.AllKeys().size(), fj.OptimizedKeys().size(), (int)lambda_rejected, (int)wrong_keys_rejected);
  return 0;
}
