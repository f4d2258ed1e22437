// JS face of the N-API addon (build/Release/cpp_utils.node).  Keeps the export name the worker
// code expects (alsBuildSubFixedFacts, dispatching on the typed-array class as upstream's
// cpp_utils/cpp_utils.js:15-19 does) and adds the step-type constants of include/ycnr_als.h.
// `ctx` is the handle returned by addon.create({...}).
'use strict';

const addon = require('../build/Release/cpp_utils');

const STEP_TYPES = Object.freeze({ byUser: 0, byItem: 1, rmseValidate: 2, rmseTest: 3 });

const gatherByClass = new Map([
  [Float32Array, addon.sAlsBuildSubFixedFacts],
  [Float64Array, addon.dAlsBuildSubFixedFacts],   // throws: the GPU path is float32 only
]);

function alsBuildSubFixedFacts(ctx, subFixedFacts, fixedFacts, indx, cols, factorsCount) {
  const impl = gatherByClass.get(subFixedFacts.constructor);
  if (impl === undefined) {
    throw new Error('invalid type!');
  }
  return impl(ctx, subFixedFacts, fixedFacts, indx, cols, factorsCount);
}

module.exports = Object.assign({}, addon, { alsBuildSubFixedFacts, STEP_TYPES });
