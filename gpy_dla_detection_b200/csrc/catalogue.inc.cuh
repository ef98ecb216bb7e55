// catalogue.inc.cuh : batched catalogue engine - included by dla_b200.cu.
struct dla_catalogue {
  int placeholder = 0;
};

extern "C" int dla_catalogue_create(const dla_model*, const dla_params*, const dla_catalogue_config*, const double*,
                                    const double*, const double*, const double*, const double*, const double*,
                                    dla_catalogue**) {
  return fail("dla_catalogue_create: not built yet");
}
extern "C" int dla_catalogue_destroy(dla_catalogue* cat) {
  delete cat;
  return 0;
}
extern "C" int dla_catalogue_process(dla_catalogue*, int, const int64_t*, const double*, const double*, const double*,
                                     const uint8_t*, const double*, const double*, dla_catalogue_outputs*) {
  return fail("dla_catalogue_process: not built yet");
}
extern "C" int dla_catalogue_stage(dla_catalogue*, int, const int64_t*, const double*, const double*, const double*,
                                   const uint8_t*, const double*, const double*) {
  return fail("dla_catalogue_stage: not built yet");
}
extern "C" int dla_catalogue_run_staged(dla_catalogue*, dla_catalogue_outputs*) {
  return fail("dla_catalogue_run_staged: not built yet");
}
extern "C" int dla_catalogue_last_timing(const dla_catalogue*, double*, double*, double*, long long*, double*) {
  return fail("dla_catalogue_last_timing: not built yet");
}
