/**
 * gpu_worker.js — worker-protocol shim over the N-API addon (spectro_napi.c).
 *
 * There is no Node.js in the build image: this file is executed in the test suite by oracle/jsmini.py
 * (tests/test_js_host.py) with the addon replaced by an object of the same three functions that calls the C ABI
 * (GPU suite) or the oracle (CPU suite).  spectro_b200/worker.py is the Python mirror of the same shim.
 * Hand the class to the reference as its worker constructor:
 *
 *     import { GpuWorker } from './gpu_worker.js'
 *     new Spectroplot({ ..., workerOrUrl: GpuWorker })     // reference lib/spectroplot.js:100-116
 *
 * The request / reply objects are exactly those of reference lib/spectroplot.js:1213-1226 and
 * lib/worker.js:140-149; one reply per request, in order; messages without `.buffer` are ignored
 * (lib/worker.js:159).  Errors from the engine reject instead of hanging the render promise.
 */
'use strict'
const addon = require('./spectro_napi.node')

// SampleView's alias table (reference lib/samples.js:22-155): unknown formats are CU8
const FORMATS = ['CU4', 'CS4', 'CU8', 'CS8', 'CU12', 'CS12', 'CU16', 'CS16', 'CU32', 'CS32', 'CU64', 'CS64', 'CF32', 'CF64']
const ALIASES = { DATA: 'CU8', COMPLEX16U: 'CU8', COMPLEX16S: 'CS8', CFILE: 'CF32', COMPLEX: 'CF32' }
function formatId(name) {
    const f = String(name).toUpperCase()
    const i = FORMATS.indexOf(ALIASES[f] || f)
    return i < 0 ? 2 : i
}

class GpuWorker {
    constructor(device = 0) {
        // a device index, or an array of indices for one worker that shards each message across several GPUs;
        // throws without an sm_100 GPU: there is no CPU fallback
        this.engine = addon.create(device)
        this.onmessage = null
        this.onerror = null
    }

    postMessage(msg /*, transfer */) {
        if (!(msg && msg.buffer)) return // lib/worker.js:159 (also swallows the {transferable} probe)
        let data
        try {
            // the reference stores the table entries into a Uint8ClampedArray (lib/worker.js:118-120): clamp to [0, 255],
            // round half to EVEN, NaN -> 0 - exactly what a Uint8ClampedArray store does, so let one do it
            const cmap = new Uint8ClampedArray(msg.cmap.length * 3)
            msg.cmap.forEach((c, i) => {
                for (let k = 0; k < 3; k++) cmap[3 * i + k] = c[k]
            })
            const r = addon.render(this.engine, {
                buffer: msg.buffer, format: formatId(msg.format), n: msg.n, width: msg.width,
                block_norm: msg.block_norm, gain: msg.gain, range: msg.range,
                windowc: Float64Array.from(msg.windowc), cmap: cmap,
                channelMode: !!msg.channelMode, waterfall: !!msg.waterfall,
            })
            data = {
                cB_hist: Array.from(r.cB_hist, Number),
                c_hist: Array.from(r.c_hist, Number),
                dBfs_min: r.dBfs_min,
                dBfs_max: r.dBfs_max,
                offset: msg.offset,
                gauge_mins: r.gauge_mins,
                gauge_maxs: r.gauge_maxs,
                gauge_amps: r.gauge_amps,
                imageData: { data: r.image },
            }
        } catch (err) {
            if (this.onerror) this.onerror(err)
            else throw err
            return
        }
        // replies are matched FIFO per worker (lib/spectroplot.js:111-115): answer asynchronously, in order
        Promise.resolve().then(() => this.onmessage && this.onmessage({ data }))
    }

    terminate() {
        addon.destroy(this.engine)
        this.engine = null
    }
}

module.exports = { GpuWorker, formatId }
