/**
 * spectroplot_headless.js — the reference's controller surface without a DOM, for Node.js.
 *
 * `new Spectroplot(options)`, `setOption`, `setOptions`, `setData`, `zoomIn / zoomOut / zoomFit`, `destroy` keep the
 * names, defaults, `constrain` parsers, return values (the render Promise) and the single-flight rule of reference
 * lib/spectroplot.js:212-306, 461-527, 1096-1285.  What a browser page gets as painted canvases, a Node caller gets as
 * buffers: the Promise resolves to { image, width, height, waterfall, cB_hist, c_hist, dBfs_min, dBfs_max,
 * gauge_mins, gauge_maxs, gauge_amps }.  Axes, ramps, themes, events and the URL loader are browser UI and are not here.
 *
 * The pure modules of the reference are INJECTED, not re-implemented:
 *
 *     const createSpectroplot = require('./spectroplot_headless')
 *     const Spectroplot = createSpectroplot({
 *         windows,                       // lib/windows.js exports            { hannWindow, ... }
 *         cmaps,                         // lib/*cmap.js exports merged       { cube1_cmap, viridis_cmap, ... }
 *         lookup,                        // lib/utils.js
 *         parseFreqRate, parseFormat,    // lib/parseFreqRate.js
 *         SampleView,                    // lib/samples.js (default export)
 *         Worker: require('./gpu_worker').GpuWorker,     // any worker-shaped constructor (lib/spectroplot.js:108)
 *     })
 *
 * Tested in this repository by executing this file with oracle/jsmini.py (tests/test_js_host.py); under Node it needs
 * nothing beyond CommonJS.
 */
'use strict'

const CB_HIST_SIZE = 1000 // 0.1 dB bins (lib/spectroplot.js:1135)

function toInt(value, otherwise) {
    return parseInt(value, 10) || otherwise
}

function createSpectroplot(deps) {
    const { windows, cmaps, lookup, parseFreqRate, parseFormat, SampleView } = deps

    class Spectroplot {
        constructor(options) {
            // option parsers, lib/spectroplot.js:238-251
            this.constrain = {
                fftN: v => toInt(v, 512),
                height: v => toInt(v, 0),
                windowF: v => lookup(windows, v) || windows.blackmanHarrisWindow,
                zoom: v => toInt(v, 1),
                gain: v => toInt(v, 0),
                range: v => toInt(v, 30),
                cmap: v => lookup(cmaps, v) || cmaps.cube1_cmap,
                ampHeight: v => toInt(v, 0),
                minmaxHeight: v => toInt(v, 0),
                histWidth: v => toInt(v, 0),
                channelMode: v => (typeof v === 'string') ? !v.toLowerCase().startsWith('i') : v,
                turnFlip: v => (typeof v === 'string') ? !v.toLowerCase().startsWith('s') : v,
            }
            const opts = Object.assign({
                fftN: 512, width: 3000, height: 512, zoom: 1, windowF: windows.blackmanHarrisWindow, gain: 6, range: 30,
                cmap: cmaps.cube1_cmap, ampHeight: 0, minmaxHeight: 20, channelMode: false, turnFlip: false,
                dbfsWidth: 60, dbfsHeight: 0, freqWidth: 40, timeHeight: 20, rampHeight: 0, rampTop: 10, rampWidth: 15,
                histWidth: 100, histLeft: 55,
                // stand-ins for what the page supplies: parent.clientWidth, window.innerHeight, navigator.hardwareConcurrency
                clientWidth: 3200, innerHeight: 3200, workerCount: 1, workerArgs: [],
            }, options)
            this.opts = opts
            this.buffer = null
            this.fileinfo = null
            this.fftN = opts.fftN
            this.width = opts.width
            this.height = opts.height
            this.zoom = opts.zoom
            this.windowF = lookup(windows, opts.windowF)
            this.gain = opts.gain
            this.range = opts.range
            this.cmap = lookup(cmaps, opts.cmap)
            this.ampHeight = opts.ampHeight
            this.minmaxHeight = opts.minmaxHeight
            this.channelMode = opts.channelMode
            this.turnFlip = opts.turnFlip
            this.histWidth = opts.histWidth
            this.inProcess = false
            this.result = null
            this.workers = []
            this.pending = []
            this.startWorkers(opts.workerOrUrl || deps.Worker, opts.workerCount, opts.workerArgs)
            if (opts.filedata) this.setData(opts.filedata)
        }

        // lib/spectroplot.js:100-130: a pool of worker-shaped objects, replies matched first-in first-out per worker
        startWorkers(WorkerCtor, count, args) {
            for (let i = 0; i < count; i++) {
                const worker = new WorkerCtor(args[i % (args.length || 1)])
                const queue = []
                worker.onmessage = msg => {
                    const waiter = queue.shift()
                    if (waiter) waiter.resolve(msg)
                }
                worker.onerror = err => {
                    const waiter = queue.shift()
                    if (waiter) waiter.reject(err)
                }
                this.workers.push(worker)
                this.pending.push(queue)
            }
        }

        renderPromise(i, message, transfer) {
            return new Promise((resolve, reject) => {
                this.pending[i].push({ resolve, reject })
                this.workers[i].postMessage(message, transfer)
            })
        }

        destroy() {
            this.workers.forEach(w => w.terminate && w.terminate())
            this.workers = []
            this.pending = []
        }

        setOption(opt, value) { // :461-464
            this[opt] = this.constrain[opt](value)
            return this.processData()
        }

        setOptions(opts) { // :471-476
            for (const opt in opts) this[opt] = this.constrain[opt](opts[opt])
            return this.processData()
        }

        setData(filedata) { // :483-511 (the URL form needs XHR and is not offered here)
            if (typeof filedata === 'string') throw new Error('setData(url) needs a browser; pass {fileBuffer, name, size, type}')
            this.fileinfo = filedata
            this.buffer = filedata.fileBuffer
            this.sampleFormat = parseFormat(filedata.name)
            const nameInfo = parseFreqRate(filedata.name)
            this.center_freq = nameInfo.freq
            this.sample_rate = nameInfo.rate
            this.sampleView = new SampleView(this.sampleFormat, null, this.sample_rate, this.center_freq)
            return this.sampleView.loadBuffer(this.buffer).then(() => {
                this.sample_rate = this.sampleView.sampleRate
                return this.processData()
            })
        }

        zoomOut() { // :513-527: half steps inside [1, 8]
            if (this.zoom <= 1) return Promise.resolve()
            this.zoom -= 0.5
            return this.processData()
        }

        zoomFit() {
            if (this.zoom == 1) return Promise.resolve()
            this.zoom = 1
            return this.processData()
        }

        zoomIn() {
            if (this.zoom >= 8) return Promise.resolve()
            this.zoom += 0.5
            return this.processData()
        }

        // lib/spectroplot.js:1096-1285 without the drawing: build one message per worker, merge the replies
        processData() {
            if (!this.buffer || !this.sampleView || !this.sampleView.buffer) return Promise.resolve() // :1097-1098
            if (this.inProcess) return this.inProcess // single flight: a second request is dropped (:1099)

            const waterfall = !!this.turnFlip
            const extra = this.opts.freqWidth + this.opts.dbfsWidth + this.histWidth
            const width = ~~((waterfall ? this.opts.innerHeight : this.opts.clientWidth) * this.zoom - extra) // :1103-1104
            this.width = width
            const n = this.fftN
            const height = n
            const win = this.windowF(n)
            const block_norm = 1.0 / win.weight // :1116
            const cmap = this.cmap
            cmap[0] = [0, 0, 0] // :1129-1130, in place like the reference
            cmap[cmap.length - 1] = [255, 255, 255]

            const sampleView = this.sampleView
            const count = this.workers.length
            const endSample = ~~(this.buffer.byteLength / sampleView.sampleWidth) // :1207
            const sliceWidth = ~~(width / count) // :1208

            const out = {
                image: new Uint8ClampedArray(4 * width * height), width: width, height: height, waterfall: waterfall,
                cB_hist: new Array(CB_HIST_SIZE).fill(0), c_hist: new Array(cmap.length).fill(0),
                dBfs_min: 0.0, dBfs_max: -200.0,
                gauge_mins: new Uint8ClampedArray(width), gauge_maxs: new Uint8ClampedArray(width), gauge_amps: new Uint8ClampedArray(width),
            }
            const renders = []
            for (let i = 0; i < count; ++i) {
                const slice = sampleView.slice(i, count, 0, endSample) // :1211
                const message = { // :1213-1226
                    block_norm: block_norm, gain: this.gain, range: this.range, cmap: cmap, n: n, windowc: win.window,
                    width: sliceWidth, offset: i * sliceWidth, buffer: slice, format: sampleView.format,
                    channelMode: this.channelMode, waterfall: waterfall,
                }
                renders.push(this.renderPromise(i, message, [slice]).then(msg => {
                    const d = msg.data
                    if (d.dBfs_min < out.dBfs_min) out.dBfs_min = d.dBfs_min // :1229-1238
                    if (d.dBfs_max > out.dBfs_max) out.dBfs_max = d.dBfs_max
                    for (let k = 0; k < CB_HIST_SIZE; ++k) out.cB_hist[k] += d.cB_hist[k]
                    for (let k = 0; k < cmap.length; ++k) out.c_hist[k] += d.c_hist[k]
                    const tile = d.imageData.data
                    const off = d.offset
                    if (waterfall) { // putImageData(tile, 0, width - sliceWidth - offset), :1244
                        out.image.set(tile, 4 * height * (width - sliceWidth - off))
                    } else { // putImageData(tile, offset, 0): a column band
                        for (let y = 0; y < height; ++y)
                            out.image.set(tile.subarray(4 * sliceWidth * y, 4 * sliceWidth * (y + 1)), 4 * (width * y + off))
                    }
                    out.gauge_mins.set(d.gauge_mins, off)
                    out.gauge_maxs.set(d.gauge_maxs, off)
                    out.gauge_amps.set(d.gauge_amps, off)
                }))
            }
            const settle = () => { this.inProcess = false }
            this.inProcess = Promise.all(renders).then(() => {
                settle()
                this.dBfs_min = out.dBfs_min
                this.dBfs_max = out.dBfs_max
                this.result = out
                return out
            }, err => {
                settle() // the reference has no .catch here and would stay stuck (lib/spectroplot.js:1277-1284)
                throw err
            })
            return this.inProcess
        }
    }
    return Spectroplot
}

module.exports = createSpectroplot
