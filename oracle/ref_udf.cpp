// =============================================================================
// oracle/ref_udf.cpp — TEST INFRASTRUCTURE, NOT THE PRODUCT (checker build only).
//
// The reference's example user-defined model, apps/udf_model/minimal.cpp, compiled as it is (included from where it
// lies) together with the reference's UdfModel dispatch (models/src/udfmodel_user.cpp).  The reference loads the hooks
// from a shared object through the un-vendored `dynlib` wrap; here the same hooks are wired into
// UnsafeUDF::Loader's function pointers directly (models/ext/udf_includes.cpp:9-50 declares them, :58-80 fills them
// from the module table).
// =============================================================================
#define DECLARE_EXPORT_UDF 1
#include <minimal.cpp>  // -I $(REF)/apps/udf_model

namespace UnsafeUDF {
std::size_t (*Loader::set_nvar_udf)() = &_set_nvar;
std::size_t (*Loader::set_nc_udf)() = &_set_nc;
void (*Loader::init_udf)(const MC::pool_type&, std::size_t, const Models::UdfModel::SelfParticle&, const Models::UdfModel::Config&) = &_init_udf;
MC::Status (*Loader::update_udf)(const MC::pool_type&, float, std::size_t, const Models::UdfModel::SelfParticle&,
                                 const Models::UdfModel::SelfContribs&, const std::size_t, const MC::LocalConcentration&) = &_update_udf;
Models::UdfModel::Config (*Loader::get_config_udf)(Kokkos::DefaultHostExecutionSpace&, std::size_t) = &_get_config_udf;
void (*Loader::division_udf)(const MC::pool_type&, std::size_t, std::size_t, const MC::DynParticlesModel<float>&,
                             const MC::DynParticlesModel<float>&) = &_division_udf;
double (*Loader::mass)(std::size_t, const MC::DynParticlesModel<float>&) = &::mass;  // the hook of minimal.cpp (inside the definition `mass` names the member)
std::vector<std::string_view> (*Loader::names)() = &_names;
std::vector<std::size_t> (*Loader::get_number)() = &_get_number;
std::vector<std::string_view> (*Loader::species)() = nullptr;
}  // namespace UnsafeUDF
