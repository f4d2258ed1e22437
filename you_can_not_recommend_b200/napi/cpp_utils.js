// JS face of the N-API addon (build/Release/cpp_utils.node).  Keeps upstream's export and its FIVE-argument
// signature — alsBuildSubFixedFacts(subFixedFacts, fixedFacts, indx, cols, factorsCount), dispatching on the
// typed-array class exactly as cpp_utils/cpp_utils.js:15-19 does — and adds the step-type constants of
// include/ycnr_als.h.  The gather runs on the process's current context (the handle of the last create()).
'use strict';

const addon = require('../build/Release/cpp_utils');

const STEP_TYPES = Object.freeze({ byUser: 0, byItem: 1, rmseValidate: 2, rmseTest: 3 });

function typeCheck(array) {
  if (array.constructor === Float64Array) return true;
  if (array.constructor === Float32Array) return false;
  throw new Error('invalid type!');
}

function alsBuildSubFixedFacts(subFixedFacts, fixedFacts, indx, cols, factorsCount) {
  return typeCheck(subFixedFacts)
    ? addon.dAlsBuildSubFixedFacts(subFixedFacts, fixedFacts, indx, cols, factorsCount)   // throws: float32 only
    : addon.sAlsBuildSubFixedFacts(subFixedFacts, fixedFacts, indx, cols, factorsCount);
}

module.exports = Object.assign({}, addon, { alsBuildSubFixedFacts, STEP_TYPES });
