"""Result container with the reference's dataset names ("and the HDF5 result schema").

The reference writes one HDF5 file per rank through HighFive (apps/core/src/dataexporter/): time series under
`records/`, one group per particle dump under `biological_model/<k>/`, scalars under `initial_parameters/`, `misc/`
and `final_result/`.  No HDF5 library is installed here, so `ResultWriter` collects the SAME datasets under the SAME
paths and stores them as one `.npz` whose keys are the dataset paths (`save_hdf5` writes real HDF5 when `h5py` is
importable).  Works with any loop object of this repository (`ParticleLoop`, and in tests the oracle).

  records/time                      append per dump                  main_exporter.cpp:110,135
  records/concentration_liquid      [t][n_comp][n_species]           main_exporter.cpp:69,132
  records/volume_liquid             [t][n_comp]                      main_exporter.cpp:79,134
  records/number_particle           [t][n_comp]  getRepartition      partial_exporter.cpp:31,88
  records/tallies                   [t][6]       event tallies       partial_exporter.cpp:41,160
  biological_model/<k>/<name>       exported property per particle   partial_exporter.cpp:107-156, post_process.cpp:159
  biological_model/<k>/spatial/<name>  its per-compartment sum
  biological_model/<k>/age_hydro, age  the two ages
  initial_parameters/*              main_exporter.cpp:32-50
  misc/n_rank, misc/species_names   main_exporter.cpp:26-29
  final_result/number_particles, events/{move,total_division,total_death,total_exit}, concentration_liquid
                                    main_exporter.cpp:153-191
"""
import numpy as np

# Model::names() (+ "mass", post_process.hpp:193-194) and Model::get_number() (None: every property, HasExportPropertiesFull)
EXPORT_NAMES = {
    "fixed_length": (["length"], [0]),                                   # fixed_length.hpp:86-98
    "monod": (["length", "mu", "mu_eff"], [0, 2, 3]),                    # monod.hpp:195-207
    "simple_acetate": (["length", "l_max", "a_p", "a_max", "a_e", "a_e_s", "a_e_a", "phi_s", "phi_a"], None),  # simple_acetate.hpp:114-122
    "udf_model": (["length"], [0]),                                      # apps/udf_model/minimal.cpp:118-127
}
SPECIES = {"fixed_length": ["S"], "monod": ["0"], "simple_acetate": ["0", "1"], "udf_model": ["0"]}


class ResultWriter:
    def __init__(self, model, n_compartments, n_species, *, number_particles, initial_weight, initial_biomass_concentration,
                 final_time, delta_time, n_rank=1, n_map=1, t_per_flow_map=0.0):
        self.model, self.n_comp, self.n_species = model, int(n_compartments), int(n_species)
        self.names, self.indices = EXPORT_NAMES[model]
        self.d = {
            "initial_parameters/number_particles": np.uint64(number_particles),
            "initial_parameters/initial_weight": np.float64(initial_weight),
            "initial_parameters/initial_biomass_concentration": np.float64(initial_biomass_concentration),
            "initial_parameters/number_compartment": np.uint64(n_compartments),
            "initial_parameters/final_time": np.float64(final_time),
            "initial_parameters/delta_time": np.float64(delta_time),
            "initial_parameters/n_map": np.uint64(n_map),
            "initial_parameters/t_per_flow_map": np.float64(t_per_flow_map),
            "misc/n_rank": np.uint32(n_rank),
            "misc/species_names": np.array(SPECIES.get(model, [str(i) for i in range(n_species)])),
        }
        self._rec = {k: [] for k in ("time", "concentration_liquid", "volume_liquid", "number_particle", "tallies")}
        self.export_counter = 0

    def update_fields(self, t, concentrations, volumes):
        """MainExporter::update_fields (main_exporter.cpp:118-151); concentrations species-fastest"""
        self._rec["time"].append(float(t))
        self._rec["concentration_liquid"].append(np.asarray(concentrations, np.float64).reshape(self.n_comp, self.n_species).copy())
        self._rec["volume_liquid"].append(np.asarray(volumes, np.float64).copy())

    def write_particle_dump(self, loop, with_age=True):
        """PostProcessing::save_particle_state -> PartialExporter (post_process.cpp:92-170): repartition, tallies and the
        particle properties of this dump (forces a compaction, like the reference)"""
        ex = loop.get_properties(None if self.indices is None else np.asarray(self.indices, np.uint64), with_age)
        self._rec["number_particle"].append(np.asarray(loop.repartition(), np.uint64).copy())
        c = loop.counters()
        self._rec["tallies"].append(np.array([c["events"][e] for e in ("NewParticle", "Exit", "Move", "Death", "Overflow", "ChangeWeight")], np.uint64))
        g = f"biological_model/{self.export_counter}/"
        if with_age:
            self.d[g + "age_hydro"] = ex["ages"][0].copy(); self.d[g + "age"] = ex["ages"][1].copy()
        for i, name in enumerate(self.names + ["mass"]):
            self.d[g + name] = ex["particle_values"][i].copy()
            self.d[g + "spatial/" + name] = ex["spatial_values"][i].copy()
        self.export_counter += 1
        return ex

    def write_final(self, loop, concentrations):
        """MainExporter::write_final (main_exporter.cpp:153-191)"""
        c = loop.counters()
        self.d["final_result/number_particles"] = np.uint64(c["n_used"] - c["n_inactive"])
        self.d["final_result/events/move"] = np.uint64(c["events"]["Move"])
        self.d["final_result/events/total_division"] = np.uint64(c["events"]["NewParticle"])
        self.d["final_result/events/total_death"] = np.uint64(c["events"]["Death"])
        self.d["final_result/events/total_exit"] = np.uint64(c["events"]["Exit"])
        self.d["final_result/concentration_liquid"] = np.asarray(concentrations, np.float64).reshape(self.n_comp, self.n_species).copy()

    def datasets(self):
        out = dict(self.d)
        for k, v in self._rec.items():
            if v:
                out["records/" + k] = np.array(v)
        return out

    def save(self, path):
        """one .npz, keys = the reference's dataset paths"""
        np.savez_compressed(path, **self.datasets())

    def save_hdf5(self, path):
        import h5py  # not installed in this image
        with h5py.File(path, "w") as f:
            for k, v in self.datasets().items():
                f.create_dataset(k, data=v.astype("S") if getattr(v, "dtype", None) is not None and v.dtype.kind == "U" else v)
